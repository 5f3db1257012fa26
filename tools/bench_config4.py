"""BASELINE config 4 on its own: 200 000 frames (224x224 uint8, generated per rank on the device, seed 4 + rank),
contiguous frame ranges over the ranks, batch 256, final NCCL gather of the (200000, 2048) fp32 features to rank 0.
The leg itself lives in bench.py (config4_leg), which also runs it under torchrun; this wrapper runs it alone.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/bench_config4.py [frames]
    python tools/bench_config4.py [frames]          # one GPU, no gather
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import phdfx  # noqa: E402


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else bench.CONFIG4_FRAMES
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = phdfx.B200Backbone(phdfx.seeded_backbone(), device=local, max_frames=bench.BATCH)
    res = bench.config4_leg(eng, dev, rank, world, total)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
