"""bench.py — headline benchmark of the B200-native MM SAM-Adapter path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): SAM ViT-L MM-adapter + Segformer head, DeLiVER-shaped RGB+LiDAR
1024x1024, bf16 storage / fp32 accumulate, batch 8 per GPU, synthetic inputs, random-init weights
(seeded, perturbed per SURVEY.md Appendix D). One step = one forward (backbone + head + upsample + argmax
-> uint8 labels) over one batch. Multi-GPU: images sharded by rank (weak scaling), no collective on the
data path; one all-gather of the per-rank 25x25 confusion matrix after the loop.

Inputs are the decoder's output: uint8 HWC frames (RGB + LiDAR projected to 3 channels, 1 byte per value); the test
pipeline's Normalize_multimodal (configs/DELIVER/...RGBLIDAR.py:70-72, norm_by_max) runs inside the patchify kernels.

The JSON line carries: value (img/s, inputs resident in HBM), e2e (same metric through the public
EncoderDecoder API with pinned-host uint8 frames, H2D + D2H inside the timed region), roofline (dominant kernel:
the tcgen05 GEMM, achieved TFLOP/s measured live with CUDA events in an instrumented pass), roofline_msda
(MSDeformAttn GB/s vs measured HBM peak), cpu_baseline (the oracle port on the host cores, 1 image),
clocks, gpu_launches.

--impl reference: times the reference's algorithm on the host CPU (the oracle port: the reference is
Python and cannot travel to the GPU box) on 1-image samples of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

VITL = dict(img_size=1024, modalities_name=["rgb", "lidar"], modalities_ch=[3, 3], init_values=1e-6,
            gamma_init_values=1e-6, patch_size=16, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4,
            drop_path_rate=0.3, drop_multimodal_path=0, conv_inplane=48, n_points=4, deform_num_heads=16,
            cffn_ratio=0.25, deform_ratio=0.5, with_cp=True, interaction_indexes=[[0, 5], [6, 11], [12, 17], [18, 23]],
            global_attn_indexes=[5, 11, 17, 23], window_size=14, arch="small", checkpoint="none")
VITL_HEAD = dict(in_channels=[1024] * 4, in_index=[0, 1, 2, 3], channels=512, dropout_ratio=0.1, num_classes=25,
                 norm_cfg=dict(type="SyncBN", requires_grad=True), align_corners=False)
WORKLOAD = "SAM ViT-L MM-adapter + Segformer head, DeLiVER-shaped RGB+LiDAR 1024x1024 inference, batch 8 per GPU"
BATCH = 8
METRIC = "img/s @1024^2 RGB+LiDAR ViT-L adapter fwd"


# Normalize_multimodal arguments of the DeLiVER RGB+LiDAR config (configs/DELIVER/...RGBLIDAR.py:70-72, 94)
PIPELINE = dict(mean=[0.485, 0.456, 0.406, 0, 0, 0], std=[0.229, 0.224, 0.225, 1, 1, 1], to_rgb=(True, True), norm_by_max=True)


def synthetic_frames(batch, hw, seed):
    """uint8 HWC frames as the image decoder hands them over: dense RGB + a 10 %-dense LiDAR projection (3 channels)."""
    H, W = (hw, hw) if isinstance(hw, int) else hw
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randint(0, 256, (batch, H, W, 3), generator=g, dtype=torch.uint8)
    aux = torch.randint(1, 256, (batch, H, W, 3), generator=g, dtype=torch.uint8)
    aux[torch.rand(batch, H, W, 3, generator=g) >= 0.1] = 0
    return rgb, aux


def host_normalise(rgb, aux):
    """The same Normalize_multimodal on the host -> the fp32 NCHW network input (the reference arm's input)."""
    m, sd_ = torch.tensor(PIPELINE["mean"]), torch.tensor(PIPELINE["std"])
    v = torch.cat((rgb.flip(-1) if PIPELINE["to_rgb"][0] else rgb, aux.flip(-1) if PIPELINE["to_rgb"][1] else aux), -1).float()
    return ((v / 255.0 - m) / sd_).permute(0, 3, 1, 2).contiguous()


def peaks():
    p = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        p.update({k: d[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d})
        p["source"] = "measured"
    except Exception:
        pass
    return p


def bench_config(world, B, graph):
    """`config` of the JSON line: identical for both arms."""
    return {"workload": WORKLOAD, "global_batch": world * B,
            "input": "uint8 HWC frames (rgb + lidar, 6.3 MB per image); Normalize_multimodal fused into the patchify kernels",
            "l2": "a 256 MB buffer is overwritten between timed steps (L2 flush, inside the timed region)",
            "parallelism": f"image-sharded x{world}, no data-path collective",
            "launch": "cuda-graph replay" if graph else "eager"}


def build_model():
    from common import build_segmentor
    return build_segmentor(VITL, VITL_HEAD, test_cfg=dict(TEST_CFG))


TEST_CFG = dict(mode="whole_dim", rescale=True, dim=(1024, 1024))


def cpu_reference_runner(sd):
    """-> (fn(x) -> labels, kind). kind "reference": the reference's OWN modules (EncoderDecoder.simple_test over the
    reference backbone + SegformerHead, MSDeformAttn through its ms_deform_attn_core_pytorch) from the staged copy
    baseline/_ref (written by __graft_entry__.build() in the build container, git-ignored, shipped with the snapshot),
    driven through tools/ref_shim.py. kind "port": the oracle restatement, only when that copy is absent."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "segmentation", "mmseg_custom")):
        os.environ["MMSAM_REFERENCE"] = ref_root
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        cwd = os.getcwd()
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                import ref_shim
                net = ref_shim.build_segmentor(VITL, VITL_HEAD, TEST_CFG)
        finally:
            os.chdir(cwd)
        net.load_state_dict(sd, strict=True)
        meta = dict(ori_shape=(1024, 1024, 3), flip=False)
        return (lambda x: net.simple_test(x, [meta] * x.shape[0], True)), "reference"
    from oracle import model as om
    return (lambda x: om.simple_test(sd, VITL, x, TEST_CFG, True)), "port"


def cpu_reference_img_per_s(sd, threads, steps, warmup, median=False):
    """The reference's CPU implementation of the path (fp32) on the host cores: 1 image of the workload per step."""
    from oracle.perturb import synthetic_batch
    torch.set_num_threads(threads)
    fn, kind = cpu_reference_runner(sd)
    x = host_normalise(*synthetic_frames(1, 1024, 1234))
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fn(x)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    spi = statistics.median(ts) if median else sum(ts) / len(ts)
    return 1.0 / spi, spi, kind


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(dev_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def instrumented_pass(eng, x):
    """Kernel-only timings of the two kernels the roofline is reported for, taken live with CUDA events on the
    launching stream. One extra single-stream forward records every GEMM / fused MSDeformAttn call and its arguments.
    The 470 GEMM launches are short (12 - 230 us) and an eager launch from Python takes ~10 us, so an event pair
    around each one in the forward would mostly time the host; instead every distinct GEMM (and MSDeformAttn)
    configuration of the forward is replayed 8 times back to back on the tensors it ran on (event pair around the 8) and weighted by how
    often the step launches it: sum(count x per-launch time) = the GEMM time of a step as the graph replay runs it."""
    from mmsam_b200 import kernels as K
    calls, mcalls = {}, {}
    og, om_, ogl = K.gemm, K.msda_fused, K.gemm_ln

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def gemm(a, w, **kw):
        r = og(a, w, **kw)
        key = (a.shape[0], w.shape[0], a.shape[1], a.stride(0), r.stride(0), kw.get("act"), kw.get("bias") is not None,
               kw.get("scale") is not None, kw.get("residual") is not None, kw.get("row_map") is not None,
               kw.get("pixel_shuffle") is not None, str(r.dtype))
        ent = calls.get(key)
        if ent is None:
            kw2 = dict(kw)
            kw2["out"] = r
            calls[key] = [1, a, w, kw2]
        else:
            ent[0] += 1
        return r

    def gemm_ln(a, w, bias, colsum, rowstat, **kw):       # LayerNorm-folded GEMMs: same kernel, counted with the rest
        r = ogl(a, w, bias, colsum, rowstat, **kw)
        key = ("ln", a.shape[0], w.shape[0], a.shape[1], a.stride(0), r.stride(0), kw.get("act"), str(r.dtype))
        ent = calls.get(key)
        if ent is None:
            kw2 = dict(kw)
            kw2["out"] = r
            calls[key] = [1, a, w, kw2, (bias, colsum, rowstat)]
        else:
            ent[0] += 1
        return r

    def msda(value, shapes, lsi, qproj, ref, n_heads, n_levels, n_points=4, out=None, geom=None):
        r = om_(value, shapes, lsi, qproj, ref, n_heads, n_levels, n_points, out, geom=geom)
        N, S, MD = value.shape
        Lq = ref.shape[0]
        # value + locations(fp32 x2) + weights(fp32) + output, bf16 value/out
        by = N * (S * MD * 2 + Lq * n_heads * n_levels * n_points * 3 * 4 + Lq * MD * 2)
        key = (N, S, MD, Lq, n_levels)
        ent = mcalls.get(key)
        if ent is None:
            # clones: the engine's buffers are reused by later layers, and the staged kernel's speed depends on the offsets
            mcalls[key] = [1, by, (value.clone(), shapes, lsi, qproj.clone(), ref, n_heads, n_levels, n_points, torch.empty_like(r)), geom]
        else:
            ent[0] += 1
        return r

    # single-stream for this pass: on the parallel graph branches (ConvNeXt towers, neck levels) two kernels share the
    # GPU and an event pair would charge each with the other's time
    saved = {k: os.environ.get(k) for k in ("MMSAM_TOWER_STREAMS", "MMSAM_NECK_STREAMS")}
    os.environ.update(MMSAM_TOWER_STREAMS="0", MMSAM_NECK_STREAMS="0")
    import mmsam_b200.engine as E
    import mmsam_b200.neck as NK
    mods = [K, E.K, NK.K]
    for m in mods:
        m.gemm, m.msda_fused, m.gemm_ln = gemm, msda, gemm_ln
    try:
        eng.segment(x)
        torch.cuda.synchronize()
    finally:
        for m in mods:
            m.gemm, m.msda_fused, m.gemm_ln = og, om_, ogl
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    out = {}
    reps, m_ms, m_work, m_n = 8, 0.0, 0.0, 0
    shapes = {}
    for cnt, by, margs, geom in mcalls.values():
        for _ in range(2):
            om_(*margs, geom=geom)
        s, e = ev(), ev()
        s.record()
        for _ in range(reps):
            om_(*margs, geom=geom)
        e.record()
        torch.cuda.synchronize()
        m_ms += cnt * s.elapsed_time(e) / reps
        m_work += cnt * by
        m_n += cnt
        shapes["injector" if margs[6] > 1 else "extractor"] = dict(bytes=by, us=s.elapsed_time(e) / reps * 1e3, count=cnt)
    out["msda"] = dict(ms=m_ms, work=m_work, launches=m_n)
    out["msda_shapes"] = shapes
    reps, g_ms, g_work, g_n = 8, 0.0, 0.0, 0
    table = []
    for key, ent in calls.items():
        cnt, a, w, kw = ent[:4]
        run = (lambda: ogl(a, w, *ent[4], **kw)) if len(ent) > 4 else (lambda: og(a, w, **kw))
        for _ in range(2):
            run()
        s, e = ev(), ev()
        s.record()
        for _ in range(reps):
            run()
        e.record()
        torch.cuda.synchronize()
        t_ms = s.elapsed_time(e) / reps
        g_ms += cnt * t_ms
        g_work += cnt * 2.0 * a.shape[0] * w.shape[0] * a.shape[1]
        g_n += cnt
        table.append(dict(M=a.shape[0], N=w.shape[0], K=a.shape[1], count=cnt, us=t_ms * 1e3, total_ms=cnt * t_ms,
                          tflops=2.0 * a.shape[0] * w.shape[0] * a.shape[1] / (t_ms * 1e-3) / 1e12,
                          epilogue=dict(act=kw.get("act"), residual=str(kw["residual"].dtype) if kw.get("residual") is not None else None,
                                        out=str(kw["out"].dtype), ln_fold=len(ent) > 4, row_map=kw.get("row_map") is not None,
                                        pixel_shuffle=kw.get("pixel_shuffle") is not None)))
    out["gemm"] = dict(ms=g_ms, work=g_work, launches=g_n, configs=len(calls))
    out["gemm_table"] = sorted(table, key=lambda r: -r["total_ms"])
    return out


def reference_kernel_times(B):
    """The reference's own CUDA forward kernel (ms_deformable_im2col_gpu_kernel, ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299)
    compiled unmodified for sm_100a into oracle/_ref/libref_msda.so (oracle/Makefile), timed on this GPU at the two shapes of
    the step, fp32 and fp16. It needs sampling locations and softmaxed weights as tensors (the reference computes them with
    ATen kernels that are NOT timed here), so this is a lower bound of the reference op's cost. -> None when the library
    is absent."""
    import ctypes
    path = os.path.join(ROOT, "oracle", "_ref", "libref_msda.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    fn = lib.ref_msda_im2col
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 8 + [ctypes.c_void_p]
    M, D, P = 16, 32, 4
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, qn, vshapes in (("injector", 4096, [(128, 128), (64, 64), (32, 32)]), ("extractor", 21504, [(64, 64)])):
        L = len(vshapes)
        sh = torch.as_tensor(vshapes, dtype=torch.long)
        S = int(sh.prod(1).sum())
        lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1])).cuda()
        shc = sh.cuda()
        g = torch.Generator(device="cuda").manual_seed(5)
        for half in (0, 1):
            dt = torch.float16 if half else torch.float32
            value = torch.randn(B, S, M, D, device="cuda", generator=g).to(dt)
            loc = torch.rand(B, qn, M, L, P, 2, device="cuda", generator=g).to(dt)
            attn = torch.softmax(torch.randn(B, qn, M, L * P, device="cuda", generator=g), -1).view(B, qn, M, L, P).to(dt)
            o = torch.empty(B, qn, M * D, device="cuda", dtype=dt)
            st = torch.cuda.current_stream().cuda_stream
            args = (value.data_ptr(), shc.data_ptr(), lsi.data_ptr(), loc.data_ptr(), attn.data_ptr(), o.data_ptr(), B, S, M, D, qn, L, P, half, st)
            for _ in range(2):
                assert fn(*args) == 0
            ts = []
            for _ in range(5):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                assert fn(*args) == 0
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            out[f"{name}_{'fp16' if half else 'fp32'}_us"] = statistics.median(ts) * 1e3
    return out


def other_configs():
    """BASELINE configs 4 (FMB 800x800 ...NEWwithcp, whole_dim_cut, 14 classes) and 5a (MUSES 1080x1920 slide, 6 crops of
    1024^2 batched into one forward, 19 classes) at full size: device-timed throughput, 1 warm-up + 3 runs (extra keys of the
    bench line; parity for both is in tests/test_model_gpu.py)."""
    from common import build_segmentor
    from oracle.perturb import synthetic_batch

    def timed(fn, n=3):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    out = {}
    cfg4 = dict(VITL, img_size=800, modalities_name=["rgb", "thermal"], conv_drop_path_rate=0.3)
    seg, _ = build_segmentor(cfg4, dict(VITL_HEAD, num_classes=14), btype="SAMAdapterbimodalMixModNewInTwinConvNEWwithcp",
                             test_cfg=dict(mode="whole_dim_cut", rescale=False, dim=(600, 800), cut_dim=(800, 600)))
    seg = seg.cuda()
    x = synthetic_batch(8, 800, kind="thermal").cuda()
    x[:, :, 600:] = 0
    ms = timed(lambda: seg.encode_decode_labels(x, (800, 800), (600, 800)))
    out["config4_fmb_800x800_whole_dim_cut"] = {"img_per_s": 8e3 / ms, "ms_per_batch_of_8": ms}
    del seg
    torch.cuda.empty_cache()
    seg, _ = build_segmentor(VITL, dict(VITL_HEAD, num_classes=19), test_cfg=dict(mode="slide", crop_size=(1024, 1024), stride=(640, 640)))
    seg = seg.cuda()
    fr = synthetic_batch(1, (1080, 1920)).cuda()
    ms = timed(lambda: seg.slide_labels(fr))
    out["config5a_muses_1080x1920_slide"] = {"frames_per_s": 1e3 / ms, "crops_per_s": 6e3 / ms, "ms_per_frame": ms}
    del seg
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the reference-kernel timing and the config-4 / 5a legs")
    ap.add_argument("--dump-gemm", default=None, help="write the per-configuration GEMM timing table (JSON) to this path")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 0)
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        seg, sd = build_model()
        del seg
        v, spi, kind = cpu_reference_img_per_s(sd, threads, args.steps, W)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": W, "ms_per_step": spi * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(args.gpus, args.batch, not args.no_graph),
            "cpu_baseline": {"value": v, "unit": "img/s", "cores": threads, "kind": kind,
                             "sample": "1 image of the batch-8 workload per step (EncoderDecoder.simple_test: backbone + head + "
                                       "resize + softmax + argmax), fp32, " + ("the reference's own modules (baseline/_ref)"
                                                                               if kind == "reference" else "oracle port")},
            "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import mmsam_b200  # noqa: F401
    from mmsam_b200 import kernels as K
    from oracle.perturb import synthetic_batch

    B = args.batch
    seg, sd = build_model()
    seg = seg.cuda()
    seg.set_input_pipeline(PIPELINE["mean"], PIPELINE["std"], to_rgb=PIPELINE["to_rgb"], norm_by_max=PIPELINE["norm_by_max"])
    eng = seg.backbone.engine(seg.decode_head)

    def rank_data(r):
        rgb, aux = synthetic_frames(B, 1024, 1234 + r * B)
        g = torch.Generator().manual_seed(99 + r)
        gt = torch.randint(0, 25, (B, 1024, 1024), generator=g, dtype=torch.uint8)
        gt[torch.rand(gt.shape, generator=g) < 0.01] = 255
        return rgb, aux, gt

    rgb_host, aux_host, gt = rank_data(rank)
    rgb_host, aux_host = rgb_host.pin_memory(), aux_host.pin_memory()
    frames = seg.u8_input(rgb_host.cuda(), aux_host.cuda())
    gt = gt.cuda()
    conf = torch.zeros((25, 25), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: uint8 frames resident in HBM; L2 flushed between steps ----------------
    run = eng.segment if args.no_graph else eng.segment_graphed   # CUDA-graph replay of the same kernels
    for _ in range(max(W, 1)):
        run(frames)
    barrier()
    clocks = ClockSampler(local)
    l0 = K.LAUNCHES + eng.graph_launches
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        flush.zero_()
        labels = run(frames)
        K.confusion(labels, gt, 25, out=conf)
    e.record()
    barrier()
    launches = K.LAUNCHES + eng.graph_launches - l0
    ms = s.elapsed_time(e)
    clk = clocks.stop()
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        confs = [torch.zeros_like(conf) for _ in range(world)]
        dist.all_gather(confs, conf)          # the only collective of the path: 25x25 int64 per rank
        conf_all = torch.stack(confs).sum(0)
    else:
        conf_all = conf
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- e2e: public API, pinned host uint8 frames -> H2D -> forward -> D2H labels ----------------
    # EncoderDecoder.stream_labels is the test-loop call: one host batch in, one host label map out, per step; the
    # copy of step i+1 overlaps the forward of step i (copy stream), every step's H2D and D2H is inside the timed region
    seg.use_cuda_graph = not args.no_graph

    def host_batches(n):
        for _ in range(n):
            yield (rgb_host, aux_host)

    for lab in seg.stream_labels(host_batches(max(min(W, 3), 1)), (1024, 1024)):
        pass
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    d2h = 0
    for lab in seg.stream_labels(host_batches(args.steps), (1024, 1024)):
        d2h = lab.numel()                    # uint8 labels in pinned host memory
    e2.record()
    barrier()
    ms2 = s2.elapsed_time(e2)
    if dist is not None:
        t = torch.tensor([ms2], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = t.item()
    e2e = world * B * args.steps / (ms2 / 1e3)
    h2d = rgb_host.numel() + aux_host.numel()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # BASELINE config 3: the sharded run must reproduce the single-process result bit for bit. Rank 0 re-runs every
    # rank's images alone (eager launches, not the graph) and compares the confusion matrix with the all-gathered one.
    single = torch.zeros_like(conf)
    for r in range(world):
        rgb_r, aux_r, gt_r = (rgb_host, aux_host, gt) if r == rank else rank_data(r)
        lab_r = eng.segment(seg.u8_input(rgb_r.cuda(), aux_r.cuda()))
        K.confusion(lab_r, gt_r.cuda(), 25, out=single)
    equal_single = bool(torch.equal(single * args.steps, conf_all))

    pk = peaks()
    inst = instrumented_pass(eng, frames)
    gm, md = inst["gemm"], inst["msda"]
    if args.dump_gemm:
        with open(args.dump_gemm, "w") as f:
            json.dump(inst["gemm_table"], f, indent=1)
    tf = gm["work"] / (gm["ms"] / 1e3) / 1e12
    gbs = md["work"] / (md["ms"] / 1e3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": bench_config(world, B, not args.no_graph),
        "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * d2h,
                "api": "EncoderDecoder.stream_labels (H2D of step i+1 overlaps the forward of step i)"},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"kernel": "gemm_bf16_kernel (tcgen05)", "bound": "tensor", "achieved": tf,
                     "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops_sustained"],
                     # ncu --set full, lin1 (32768x4096x1024, GELU) launch: dram__bytes_read + dram__bytes_write
                     # (profiles/r02_gemm_ncu_summary.txt); algorithmic bytes of that launch: 343 MB
                     "traffic": 289.3e6, "traffic_of": "lin1 GEMM launch (M=32768 N=4096 K=1024): dram__bytes_read 75.6 MB + dram__bytes_write "
                                                        "213.7 MB of one ncu --set full capture of this kernel (profiles/r02_gemm_ncu_summary.txt, "
                                                        "launch 0); algorithmic 343 MB (A 67 + W 8 + out 268); not re-measured by this run",
                     "peak_source": pk["source"] + " (sustained: kernel timed inside a long step)",
                     "launches_per_step": gm["launches"], "distinct_configs": gm["configs"], "ms_per_step": gm["ms"],
                     "timing": "every distinct GEMM configuration of the step replayed 8x back to back (CUDA events), weighted by its launch count",
                     "share_of_step": gm["ms"] / (ms / args.steps)},
        "roofline_msda": {"kernel": "msda_fused_coop_kernel (injector, L1 gathers) + msda_staged2_kernel (extractor, shared-memory gathers)", "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"],
                          "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                          # ncu --set full: dram read + write per launch, injector (L1-gather kernel,
                          # profiles/r01_msda_ncu_summary.txt) 280.1 MB / extractor (staged kernel,
                          # profiles/r02_msda_staged2_ncu_summary.txt) 305.6 MB, averaged over the step's 4 + 6 launches
                          "traffic": (4 * 280.1e6 + 6 * 305.6e6) / 10, "traffic_of": "average msda launch of the step (not re-measured by this run)",
                          "ceiling": "SM side, not HBM: the gathered bytes are 8.2x the algorithmic bytes. Injector: L1 gathers peak at "
                                     "11.5 TB/s chip-wide (tools/micro/gather_bench.cu) -> 22 % of HBM peak on algorithmic bytes; extractor: "
                                     "shared-memory gathers, bound by instruction issue (474 warp instructions per 8 items, half of them "
                                     "bf16 -> fp32 unpacks; profiles/r02_msda_notes.md)",
                          "launches_per_step": md["launches"],
                          "ms_per_step": md["ms"], "peak_source": pk["source"]},
        "miou_check": {"pixels": int(conf_all.sum().item()), "equal_single": equal_single,
                       "equal_single_of": f"all-gathered confusion matrix of {world} rank(s) x {args.steps} steps == {args.steps} x one single-process "
                                          f"eager pass over the same {world * B} images"},
    }
    if not args.no_extras:
        rk = reference_kernel_times(B)
        if rk is not None:
            # same algorithmic bytes (bf16 value / output convention of SURVEY 8(d)) over the reference kernel's time
            mc = inst["msda_shapes"]
            rk["reference_kernel_gbs_fp32"] = (4 * mc["injector"]["bytes"] + 6 * mc["extractor"]["bytes"]) / \
                ((4 * rk["injector_fp32_us"] + 6 * rk["extractor_fp32_us"]) * 1e-6) / 1e9
            rk["reference_kernel_gbs_fp16"] = (4 * mc["injector"]["bytes"] + 6 * mc["extractor"]["bytes"]) / \
                ((4 * rk["injector_fp16_us"] + 6 * rk["extractor_fp16_us"]) * 1e-6) / 1e9
            rk["ours_us"] = {k: v["us"] for k, v in mc.items()}
            rk["note"] = ("the reference's kernel needs sampling locations + softmaxed weights materialised by separate ATen kernels "
                          "(not timed); ours computes them in-kernel from the raw projection")
            line["roofline_msda"]["reference_kernel"] = rk
        del eng
        seg.backbone.invalidate()
        torch.cuda.empty_cache()
        line["other_configs"] = other_configs()
    if not args.no_cpu_baseline:
        v, spi, kind = cpu_reference_img_per_s(sd, threads, 3, 1, median=True)
        line["cpu_baseline"] = {"value": v, "unit": "img/s", "cores": threads, "kind": kind,
                                "sample": f"1 image of the batch-{B} workload per run, 1 warm-up + 3 timed, median {spi:.1f} s, fp32, "
                                          + ("the reference's own modules (baseline/_ref)" if kind == "reference" else "oracle port")}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
