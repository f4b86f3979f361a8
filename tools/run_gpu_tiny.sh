# last seconds of the round's GPU budget: smoke() + the reworked graph test in ONE process, then the bench line
python - <<'P'
import sys, time
t0 = time.time()
sys.path.insert(0, ".")
import __graft_entry__ as g
try:
    g.smoke()
except Exception as e:
    print("SMOKE FAILED:", repr(e)[:500])
print("t=%.0fs" % (time.time() - t0), flush=True)
import pytest
pytest.main(["tests/test_gpu_graph.py", "-m", "gpu", "-q", "-s", "-x"])
print("t=%.0fs" % (time.time() - t0), flush=True)
P
timeout 60 python bench.py --no-cpu-baseline 2>/dev/null
