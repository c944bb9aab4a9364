"""Developer tool: where does the host time of one frame go?  (cProfile over bench.py's harness)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

h = bench.Harness("float16", torch.device("cuda", 0))
frames = bench.make_frames_u8(bench.N_FRAMES, h.cfg).pin_memory()   # decoded uint8 frames: the e2e arm's input
h.prime(h.eng.preprocessor()(frames[0]))
for i in range(40):
    h.step(frames[i % 32])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(64):
    h.step(frames[i % 32])
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 64 * 1e3)
pr = cProfile.Profile()
pr.enable()
for i in range(64):
    h.step(frames[i % 32])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
