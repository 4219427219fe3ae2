"""Host-side profile (cProfile) of the C4 training chunk at N rays: where the enqueue time goes."""
import sys, cProfile, pstats, io, torch
sys.argv = [sys.argv[0]] + sys.argv[1:]
src = open("scripts/time_c4.py").read().split("for _ in range(3): step()")[0]
exec(src)
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5): step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
ps = pstats.Stats(pr, stream=s).sort_stats("cumulative")
ps.print_stats(70)
print(s.getvalue()[:14000])
