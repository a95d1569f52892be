import sys, os
sys.path.insert(0, "/root/repo")
import torch
from actalker_b200 import SS2D_cond_v10, mamba_layer as ml
torch.manual_seed(0)
dtype = torch.float32
layer = SS2D_cond_v10(d_model=320, d_cond=1024, cond_size=32, dropout=0.1, d_state=16, size=72, scan_type="sweep", num_direction=2).eval()
if os.environ.get("DBG_TRAINED"):
    with torch.no_grad():
        for u in (layer.audio_unit, layer.exp_unit):
            u.A_logs.add_(0.5 * torch.randn_like(u.A_logs))
layer = layer.cuda()
B, L = int(os.environ.get("DBG_B", "25")), 5184
x = torch.randn(B, L, 320, device="cuda"); idm = torch.randn(B, 1, 1024, device="cuda"); cd = torch.randn(B, 33, 1024, device="cuda")
ones = torch.ones(1, 1, 576, 576, device="cuda")
with torch.no_grad():
    proj = layer.project_inputs(x, idm, cd, [ones, ones])
    xz1, xz2, t1, t2, m1, m2 = proj
    def run(chain):
        ml.SCAN_SEGMENTS, ml.SCAN_CHAIN = 1, chain
        res = ml._scan_branches([layer.audio_unit, layer.exp_unit], [xz1, xz2], [t1, t2], [m1.idx, m2.idx], [m1.n_sel, m2.n_sel], B, L, idx64s=[m1.idx64, m2.idx64])
        torch.cuda.synchronize()
        return [r[0].clone() for r in res]
    ref = run(0)
    for chain in (2, 8):
        got = run(chain)
        for br in range(2):
            d = (got[br] - ref[br])
            bad = ~(d == 0) | torch.isnan(got[br])
            dd = torch.nan_to_num(d, nan=0.0).abs()
            print("chain", chain, "branch", br, "bad elems", int(bad.sum()), "nan", int(torch.isnan(got[br]).sum()), "max abs diff", float(dd.max()), "ref absmax", float(ref[br].abs().max()))
            nanm = torch.isnan(got[br])
            if nanm.any():
                ni = nanm.nonzero()
                print("  NaN dirs", ni[:,0].unique().tolist(), "batches", ni[:,1].unique().tolist()[:8], "rows", int(ni[:,2].min()), int(ni[:,2].max()), "chan", int(ni[:,3].min()), int(ni[:,3].max()))
            if bad.any():
                idx = bad.nonzero()
                print("  dirs", idx[:, 0].unique().tolist(), "batches", idx[:, 1].unique().tolist()[:30])
                rows = idx[:, 2].unique()
                print("  rows min/max/count", int(rows.min()), int(rows.max()), rows.numel(), "chan blocks", (idx[:, 3] // 64).unique().tolist())
    print("---- structure")
    ref2 = run(0)
    print("ref deterministic:", all(torch.equal(a, b) for a, b in zip(ref, ref2)))
    got = run(int(os.environ.get("DBG_CHAIN", "2")))
    for br in range(2):
        for k in range(2):
            d = (got[br][k] - ref[br][k]).abs()          # (B, L, D)
            per_row = d.amax(dim=(0, 2))                  # (L,)
            nz = (per_row > 0).nonzero().view(-1)
            print("br", br, "k", k, "rows with any diff:", nz.numel(), "first", nz[:5].tolist(), "last", nz[-5:].tolist(),
                  "max per-row diff at", int(per_row.argmax()), float(per_row.max()))
            big = (per_row > 1e-3).nonzero().view(-1)
            print("    rows with diff>1e-3:", big.numel(), big[:8].tolist(), big[-8:].tolist() if big.numel() else [])
            per_b = d.amax(dim=(1, 2)); per_c = d.amax(dim=(0, 1)).view(-1, 64).amax(1)
            print("    per batch", [round(float(v), 4) for v in per_b], "per chan block", [round(float(v), 4) for v in per_c])
