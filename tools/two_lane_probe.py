"""Experiment: refine a batch as two half-batches on two streams (two engines) vs one launch chain."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from catre_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = synth.load_weights()
b = synth.make_batch(B, 1024, seed=2).to("cuda")
full = engine.Engine(1024, B, "f16x3", 0); full.load_weights(w)
per = B // lanes
engs = [engine.Engine(1024, per, "f16x3", 0) for _ in range(lanes)]
for e in engs: e.load_weights(w)
streams = [torch.cuda.Stream() for _ in range(lanes)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run_full():
    return full.refine(b.pcl, b.prior, b.init_pose, b.init_scale, b.K, 4)

def run_lanes():
    outs = []
    cur = torch.cuda.current_stream()
    for i, (e, s) in enumerate(zip(engs, streams)):
        s.wait_stream(cur)
        sl = slice(i * per, (i + 1) * per)
        with torch.cuda.stream(s):
            outs.append(e.refine(b.pcl[sl], b.prior[sl], b.init_pose[sl], b.init_scale[sl], b.K[sl], 4))
    for s in streams: cur.wait_stream(s)
    return torch.cat([o[0] for o in outs], 1), torch.cat([o[1] for o in outs], 1)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    ts.sort()
    return ts[len(ts) // 2]

p0, s0 = run_full(); p1, s1 = run_lanes(); torch.cuda.synchronize()
print("bit-exact:", torch.equal(p0, p1) and torch.equal(s0, s1))
print(f"B={B} one chain {timeit(run_full):.3f} ms ; {lanes} lanes {timeit(run_lanes):.3f} ms")
