"""cProfile of the host side of one config step (profiles/configs.py): python profiles/host_profile.py pose_step"""
import cProfile, pstats, sys, time
import torch
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import configs
name = sys.argv[1]
dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
step, info = configs.BUILDERS[name](dev)
for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
torch.cuda.synchronize()
print(f"{name}: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms/step wall")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(30)
