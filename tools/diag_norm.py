import sys, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import helpers
from helpers import port, syn
import fastposecnn_b200 as fp
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.randn(1, 2, 512, 512, generator=g).to(dev)
nrm = torch.norm(x, dim=1)
xs, ys = x[:, 0], x[:, 1]
xd, yd = xs.double(), ys.double()
f = lambda t: t.float()
v = {"fma(x,x,y*y)": torch.sqrt(f(xd * xd + f(ys * ys).double())),
     "fma(y,y,x*x)": torch.sqrt(f(yd * yd + f(xs * xs).double())),
     "x*x+y*y": torch.sqrt(f(xs * xs) + f(ys * ys)),
     "sqrt in double": f(torch.sqrt(xd * xd + yd * yd))}
for k, t in v.items():
    print("torch.norm ==", k, ":", float((t == nrm).float().mean()))
ref = port.normalize(x, 1)
mask = torch.zeros(1, 7, 512, 512, device=dev); mask[:, 1] = 5
logits = {"mask": mask, "quaternion": torch.randn(1, 24, 512, 512, device=dev), "scales": torch.rand(1, 18, 512, 512, device=dev),
          "xy": torch.cat([x, torch.zeros(1, 10, 512, 512, device=dev)], 1).contiguous(), "z": torch.zeros(1, 6, 512, 512, device=dev)}
out = fp.class_compression(logits, 7)["xy"]
print("kernel == torch normalize:", float((out == ref).float().mean()))
for k, t in v.items():
    print("kernel == x /", k, ":", float((out == x / t.unsqueeze(1)).float().mean()), "  torch normalize == that:", float((ref == x / t.unsqueeze(1)).float().mean()))
