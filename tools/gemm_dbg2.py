"""GEMM perf-debug sweep: MMSAM_GEMM_DBG bits (1 skip stores, 4 skip epilogue, 8 skip MMA issue, 16 skip TMA loads)."""
import math, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, ROOT)
    import mmsam_b200
    from mmsam_b200 import kernels as K
    shapes = ((32768, 3072, 1024, None, False), (32768, 4096, 1024, "gelu", False), (32768, 1024, 4096, None, True),
              (32768, 1536, 384, "gelu", False), (32768, 384, 1536, None, True))
    if os.environ.get("SHAPES"):   # "M,N,K,act,res;..."
        shapes = []
        for t in os.environ["SHAPES"].split(";"):
            m, n, k, act, res = t.split(",")
            shapes.append((int(m), int(n), int(k), None if act in ("", "none") else act, res == "1"))
    bns = [int(x) for x in os.environ.get("BNS", "256,128").split(",")]
    for (M, N, Kd, act, res) in shapes:
        a = torch.randn(M, Kd, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, Kd, device="cuda") / math.sqrt(Kd)).to(torch.bfloat16)
        b = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if res else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for bn in bns:
            for _ in range(3):
                K.gemm(a, w, bias=b, act=act, residual=r, out=out, block_n=bn)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                K.gemm(a, w, bias=b, act=act, residual=r, out=out, block_n=bn)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            print(f"  dbg={os.environ.get('MMSAM_GEMM_DBG','0')} M={M} N={N} K={Kd} act={act} res={res} bn={bn}: {ms*1e3:.0f} us {2.0*M*N*Kd/ms/1e9:.0f} TFLOP/s", flush=True)
else:
    for d in sys.argv[1:] or ["0"]:
        env = dict(os.environ, MMSAM_GEMM_DBG=d)
        subprocess.run([sys.executable, __file__, "child"], env=env)
