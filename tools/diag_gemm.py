"""GPU diagnostic for the tcgen05 GEMM/conv family: runs each case in a subprocess (a trap poisons the CUDA
context) and prints error structure instead of a bare assert.  Usage: python tools/diag_gemm.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["lin_fwd_128x128x32", "lin_fwd_128x128x256", "lin_fwd_300x256x256", "lin_dgrad_128x128x128", "lin_dgrad_300x256x256",
         "lin_wgrad_128x128x128", "lin_wgrad_1000x256x256", "conv3_s1", "conv3_s2", "conv1_s2", "conv3_s1_dgrad", "conv3_s2_dgrad",
         "conv3_s1_wgrad", "conv3_s2_wgrad", "lin_fwd_wide"]


def structure(name, got, ref):
    import torch
    err = (got - ref).abs()
    rel = float(err.max() / ref.abs().max().clamp_min(1e-20))
    print(f"[{name}] rel err {rel:.3e}; max|ref| {float(ref.abs().max()):.3e}; nan={bool(torch.isnan(got).any())}")
    if rel > 2e-3:
        g2 = got.reshape(-1, got.shape[-1])
        r2 = ref.reshape(-1, ref.shape[-1])
        bad = ((g2 - r2).abs() > 2e-3 * r2.abs().max())
        print(f"   bad fraction {float(bad.float().mean()):.4f}; bad rows {int(bad.any(1).sum())}/{bad.shape[0]}; "
              f"bad cols {int(bad.any(0).sum())}/{bad.shape[1]}")
        rows = bad.any(1).nonzero().flatten()[:12].tolist()
        cols = bad.any(0).nonzero().flatten()[:12].tolist()
        print(f"   first bad rows {rows}; first bad cols {cols}")
        print("   got[0,:8] ", [round(float(v), 4) for v in g2[0, :8]])
        print("   ref[0,:8] ", [round(float(v), 4) for v in r2[0, :8]])
        ratio = (g2[:4, :4] / r2[:4, :4])
        print("   ratio[:4,:4]", [[round(float(v), 3) for v in row] for row in ratio])
        zero = float((g2 == 0).float().mean())
        print(f"   zeros in output: {zero:.3f}")
    return rel


def run_case(case):
    import torch
    import torch.nn.functional as F
    from monodetr_b200 import tc
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(1)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    if case.startswith("lin_"):
        if case == "lin_fwd_wide":
            M, N, K = 40000, 512, 256
            kind = "fwd"
        else:
            _, kind, dims = case.split("_")
            M, N, K = (int(v) for v in dims.split("x"))
        x, w, dy = rn(M, K), rn(N, K) / K ** 0.5, rn(M, N)
        if kind == "fwd":
            structure(case, tc.linear_forward(x, w), x @ w.t())
            # structured input: identity-like weight to expose layout errors
            if N == K or True:
                w2 = torch.zeros(N, K, device="cuda"); idx = torch.arange(min(N, K)); w2[idx, idx] = 1.0
                structure(case + "/eye", tc.linear_forward(x, w2), x @ w2.t())
        elif kind == "dgrad":
            structure(case, tc.linear_dgrad(dy, w), dy @ w)
        else:
            structure(case, tc.linear_wgrad(dy, x), dy.t() @ x)
        return
    cfgs = {"conv3_s1": (2, 24, 80, 64, 128, 3, 1, 1), "conv3_s2": (2, 24, 80, 64, 128, 3, 2, 1), "conv1_s2": (2, 24, 80, 64, 128, 1, 2, 0)}
    base = "_".join(case.split("_")[:2])
    B, H, W, Cin, Cout, k, s, pad = cfgs[base]
    x = rn(B, Cin, H, W); w = rn(Cout, Cin, k, k) / (Cin * k * k) ** 0.5
    xn = x.permute(0, 2, 3, 1).contiguous()
    wp = tc.pack_weight(w)
    ref = F.conv2d(x, w, None, stride=s, padding=pad)
    if case.endswith("dgrad") or case.endswith("wgrad"):
        dy = rn(*ref.shape)
        xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
        F.conv2d(xr, wr, None, stride=s, padding=pad).backward(dy)
        dyn = dy.permute(0, 2, 3, 1).contiguous()
        if case.endswith("dgrad"):
            structure(case, tc.conv2d_dgrad(dyn, wp, xn.shape, None, None, k, k, s, pad), xr.grad.permute(0, 2, 3, 1))
        else:
            structure(case, tc.unpack_wgrad(tc.conv2d_wgrad(dyn, xn, None, k, k, s, pad), k, k), wr.grad)
    else:
        structure(case, tc.conv2d_forward(xn, wp, None, None, k, k, s, pad), ref.permute(0, 2, 3, 1))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
        import torch
        torch.cuda.synchronize()
    else:
        cases = sys.argv[1:] or CASES
        for c in cases:
            try:
                r = subprocess.run([sys.executable, __file__, "--one", c], capture_output=True, text=True, timeout=120)
                out = (r.stdout + ("\n" + r.stderr[-1500:] if r.returncode else "")).strip()
                print(out if out else f"[{c}] no output rc={r.returncode}")
                if r.returncode:
                    print(f"[{c}] FAILED rc={r.returncode}")
            except subprocess.TimeoutExpired:
                print(f"[{c}] TIMEOUT")
            sys.stdout.flush()
