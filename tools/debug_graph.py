import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tramba_b200
from tramba_b200.model import build_model, tsod_loss
dev = torch.device("cuda:0")
dims, bs, bench = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3]))
torch.backends.cudnn.benchmark = bench
model = build_model(img_size=384, dims=dims).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)
x = torch.randn(bs, 3, 384, 384, device=dev); y = (torch.rand(bs, 1, 384, 384, device=dev) > 0.5).float()
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    loss = tsod_loss(outs, y); loss.backward(); opt.step(); return loss
for _ in range(3): step()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
opt.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(x)
        loss = tsod_loss(outs, y); loss.backward(); opt.step()
    g.replay(); torch.cuda.synchronize(); print("capture ok, loss", loss.item())
except Exception:
    print("CAPTURE FAILED", dims, bs, bench); traceback.print_exc(limit=3)
