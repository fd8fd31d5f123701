"""Latency of one halo exchange per level (device time, CUDA events on the context's stream), one process per GPU:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/exchange_bench.py [R] [reps]
Run with SMF_P2P=0 for the NCCL send/recv transport."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aspect_b200 import StokesMatrixFreeHandler, benchmarks, synthetic  # noqa: E402

REPS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    hier = synthetic.BoxHierarchy(3, R, None, REPS[world], rank, world)
    eta_q, rhs, ps = benchmarks.nsinker_problem(hier, 8, 1e4, "harmonic average")
    h = StokesMatrixFreeHandler(3, device=torch.cuda.current_device())
    h.setup_dofs(hier, "zero")
    h.assemble(eta_q, "harmonic average", ps)
    rows = []
    for level in range(R + 1):
        for family in (0, 1):
            v = h.initialize_dof_vector(family, level)
            v.zero()
            for _ in range(20):
                h.exchange_add(v, family, level)
            h.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                h.exchange_add(v, family, level)
            h.synchronize()
            dt = (time.perf_counter() - t0) / reps
            rows.append((level, "up"[family], dt * 1e6))
            v.free()
    if rank == 0:
        print(f"exchange_bench world={world} R={R} transport={h.exchange_transport()!r} (us per exchange, wall clock over {reps} back-to-back calls)")
        for level, fam, us in rows:
            print(f"  level {level} {fam}: {us:8.2f} us")
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
