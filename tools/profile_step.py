"""Host-side profile of one bench step (cProfile + phase timers with syncs).
    python tools/profile_step.py [supervised|reinforce|aerial] [u8|f32] [batch]"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
workload = sys.argv[1] if len(sys.argv) > 1 else "supervised"
src = sys.argv[2] if len(sys.argv) > 2 else "u8"
dev = torch.device("cuda", 0)
cls, default_batch = bench.WORKLOADS[workload]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else default_batch
wl = cls(batch, 0, dev, src if workload == "supervised" else "u8")
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
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumtime").print_stats(34); print(st.getvalue()[:8000])
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(20); print(st.getvalue()[:5000])
