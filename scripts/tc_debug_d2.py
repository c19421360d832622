import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import mentflow_b200 as mf
from mentflow_b200 import ops
from mfb_testutil import oracle_from_generator
torch.manual_seed(2)
d = 2
gen = mf.generate.NSFGenerator(d)
with torch.no_grad():
    for p in gen.parameters():
        p.mul_(2.0)
ref = oracle_from_generator(gen)
gen = gen.to("cuda")
z = torch.randn(20000, d)
with torch.no_grad():
    steps64 = ref.forward_steps(z.double())
    for flag in (True, False):
        ops.NSF_USE_TENSOR_CORES = flag
        # single layer errors: feed the float64 layer input (rounded to fp32) to each layer
        print("tensor cores", flag)
        for t in range(gen.transforms):
            vin = steps64[t].float().cuda()
            packed = gen.packed_parameters()
            images = ops.nsf_tc_images(packed, gen._orders, 64, 3, 20) if flag else None
            y, lq = ops.nsf_layer_forward(vin, packed[t], gen._orders[t], 64, 3, 20, None, True, True,
                                          image=None if images is None else images[t])
            # oracle single layer in float64 from the same rounded input
            yr, ladj = ref.layers[t].call_and_ladj(vin.cpu().double())
            if yr is None:
                break
            ey = (y.cpu().double() - yr).abs()
            base = -0.5 * (vin.cpu().double() ** 2).sum(1) - d * 0.9189385332046727
            el = (lq.cpu().double() - (base - ladj)).abs()
            print(f"  layer {t} order {gen._orders[t]}: |dy| per feature median {[float(ey[:, i].median()) for i in range(d)]} max {[float(ey[:, i].max()) for i in range(d)]}  |dlogq| median {float(el.median()):.2e} max {float(el.max()):.2e}")
