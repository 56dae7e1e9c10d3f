"""roitr_b200: the RoITr forward hot path on B200 (sm_100a). See DESIGN.md."""
import os

# A step is one CUDA graph whose branches (main stream, sampling / search / global-transformer lanes, 8 head streams) are
# meant to run concurrently. The driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware work queues (default 8);
# branches that share a queue execute in capture order and a blocked node stalls everything behind it (measured with
# scripts/timeline.py: the search lane started 5.8 ms late). Must be set before the CUDA context is created.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
