"""Single-GPU reproduction of tests/test_gpu_ddp.py's worker (NCCL world of one rank): prints the traceback."""
import os, sys, queue
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_ddp as T
overlap = len(sys.argv) > 1 and sys.argv[1] == "1"
q = queue.Queue()
T._worker_body(0, 1, T._free_port(), q, overlap)
print("result", q.get())
