"""Host-side profile of one supervised bench step (cProfile + phase timers with syncs)."""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
src = sys.argv[1] if len(sys.argv) > 1 else "f32"
dev = torch.device("cuda", 0)
wl = bench.SupervisedWorkload(256, 0, dev, src)
wl.to_device()
for s in range(3):
    wl.run(s)
torch.cuda.synchronize()
for s in range(3, 6):
    t0 = time.perf_counter(); wl.run(s); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"step {s}: host {1e3*(t1-t0):.2f} ms, +sync {1e3*(t2-t1):.2f} ms")
pr = cProfile.Profile(); pr.enable()
for s in range(6, 11):
    wl.run(s)
torch.cuda.synchronize()
pr.disable()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumtime").print_stats(28); print(st.getvalue()[:6000])
