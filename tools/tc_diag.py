import sys, torch
sys.path.insert(0, '/root/repo')
from tests.util import load_golden, build_cuda_nef
DEV='cuda'
g = load_golden("trace_delta_permuto_ray")
nef = build_cuda_nef(g, DEV)
# round weights to fp16-representable values so that only activation rounding differs
with torch.no_grad():
    for n,p in nef.named_parameters():
        if 'decoder' in n: p.copy_(p.half().float())
for M in (128, 4096, 65536):
    gen = torch.Generator().manual_seed(M)
    coords = (torch.rand(M, 1, 3, generator=gen) * 2 - 1).to(DEV)
    ray_d = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1).to(DEV)
    chans = {'density', 'rgb', 'semantics', 'inst_embedding'}
    res = {}; gws=None
    for prec in ('fp32', 'fp16'):
        nef.decoder_precision = prec
        nef.zero_grad(set_to_none=True)
        ct, dt = coords.clone().requires_grad_(True), ray_d.clone().requires_grad_(True)
        out = nef(coords=ct, ray_d=dt, channels=chans)
        if gws is None:
            gws = {c: torch.randn(out[c].shape, generator=gen).to(DEV) * 1e-3 for c in chans}
        sum((out[c] * gws[c]).sum() for c in chans).backward()
        res[prec] = ({c: out[c].detach() for c in chans}, {k: p.grad.clone() for k, p in nef.named_parameters()}, ct.grad, dt.grad)
    print("==== M", M)
    for c in chans:
        a,b = res['fp16'][0][c], res['fp32'][0][c]
        print(f"out {c:16s} rel l2 {float((a-b).norm()/b.norm()):.2e} max {float((a-b).abs().max()/b.abs().max()):.2e}")
    for k in res['fp32'][1]:
        a,b = res['fp16'][1][k], res['fp32'][1][k]
        print(f"grad {k:40s} rel l2 {float((a-b).norm()/b.norm()):.2e} max {float((a-b).abs().max()/b.abs().max()):.2e}")
    for nm,i in (("coords",2),("ray_d",3)):
        a,b = res['fp16'][i], res['fp32'][i]
        print(f"grad {nm:40s} rel l2 {float((a-b).norm()/b.norm()):.2e} max {float((a-b).abs().max()/b.abs().max()):.2e}")
