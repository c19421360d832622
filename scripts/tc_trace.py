import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import mentflow_b200 as mf
from mentflow_b200 import _lib
torch.manual_seed(0)
gen = mf.generate.NSFGenerator(6)
with torch.no_grad():
    for p in gen.parameters(): p.mul_(3.0)
gen = gen.to("cuda")
z = torch.randn(1_000_000, 6, device="cuda")
with torch.no_grad():
    for _ in range(2): gen.forward_and_log_prob(z)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros(4 * 1024, dtype=np.int64)
rc = lib.mfb_debug_copy_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0
names = {1: "wait>", 2: "<wait", 3: "reqL1", 4: "reqH0", 5: "reqH1", 6: "reqH2"}
for s in range(5): names[10 + s] = f"ld{s}"; names[20 + s] = f"spl{s}"
ev = []
for w in range(4):
    a = buf[w * 1024:(w + 1) * 1024]
    a = a[a != 0]
    ev.append([(int(x) >> 48, int(x) & 0xFFFFFFFFFFFF) for x in a])
t0 = min(e[0][1] for e in ev if e)
# print warp 0 timeline for events 200..330 with deltas; and skew of the other warps at same index
n = min(len(e) for e in ev)
print("events per warp", [len(e) for e in ev])
for i in range(300, min(n, 420)):
    ids = [ev[w][i][0] for w in range(4)]
    ts = [ev[w][i][1] - t0 for w in range(4)]
    d = ts[0] - (ev[0][i - 1][1] - t0)
    print(f"{i:4d} {names.get(ids[0], ids[0]):7s} t={ts[0]:9d} dt={d:6d}  skew vs w0: {[t - ts[0] for t in ts[1:]]} ids_same={len(set(ids)) == 1}")
