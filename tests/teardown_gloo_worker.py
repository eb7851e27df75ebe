"""Worker for tests/test_slabs_gloo.py::test_two_phase_teardown_order_over_gloo: every rank drives slabs.close_ring with a
stand-in solver that only records when its mappings were dropped and when its memory was freed."""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gloo_worker import load_by_path  # noqa: E402


class RecordingSolver:
    def __init__(self):
        self.events = {}

    def ipc_disconnect(self):
        time.sleep(random.random() * 0.2)  # ranks reach the phases at different times
        self.events["disconnect"] = time.time()

    def close(self):
        time.sleep(random.random() * 0.2)
        self.events["close"] = time.time()


def main():
    out = sys.argv[1]
    import torch.distributed as dist
    slabs = load_by_path("lbm_slabs", "lbm-graphcore_b200", "slabs.py")
    dist.init_process_group(backend="gloo", init_method="env://")
    rank = dist.get_rank()
    random.seed(rank)
    s = RecordingSolver()
    time.sleep(random.random() * 0.2)
    s.events["still_running_until"] = time.time()  # the last moment this rank may still push halos into its neighbours
    slabs.close_ring(s, dist)
    slabs.barrier_before_reupload(dist)
    with open(os.path.join(out, f"rank{rank}.json"), "w") as f:
        json.dump(s.events, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
