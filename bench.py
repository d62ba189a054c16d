#!/usr/bin/env python
"""bench.py -- headline benchmark of the instantvnr hot path on B200.

Workload (BASELINE.json configs[1], "vnr_cmd_render-equivalent"): synthetic 256^3 volume, the
example-model.json network (trained here for a few hundred steps, untimed), one 1024^2 frame per
step with macrocell space skipping, rendering mode 5 (sample streaming).  A "step" is one frame.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = neural samples decoded per second over the K timed frames
with everything resident in HBM (device-timed with CUDA events on the renderer's stream, no frame
download); `e2e` = the same metric through the public C-ABI call sequence vnr_render + vnr_map_frame
with the frame copied back to pinned host memory every step.  `roofline` describes the dominant kernel
(fused hash-grid + tcgen05 MLP decode): algorithmic bytes per decoded sample (1024 B gathered + 16 B
sample record + 4 B value) / its CUDA-event time, against the measured HBM copy bandwidth.
`cpu_baseline` times the CPU oracle (scalar port, OpenMP) on a bounded sub-frame of the same scene.
N > 1: image strips are dealt round-robin to the ranks (tile-parallel, strong scaling) and gathered to
rank 0 over NCCL inside the timed region.
--impl reference: the reference's own decode (tiny-cuda-nn built from /root/reference, oracle/_ref) on
sample batches of the same size on the same GPU -- the reference has no CPU path and its marcher cannot
be built offline -- falling back to the CPU oracle port when that library is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "neural_samples_per_sec_render_1024"
UNIT = "samples/s"
BYTES_PER_SAMPLE = 1024 + 16 + 4          # gather (8 levels x 8 corners x 16 B) + sample record + value


_REAL_STDOUT = None


def isolate_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner): route fd 1 to
    stderr for the whole run and keep the original stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML in-process every 5 ms (no start-up latency, so even a
    0.2 s timed region is covered), `nvidia-smi -lms` as the fallback when pynvml is unusable."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.nvml = None; self.handle = None; self.samples = []; self.bits = 0; self.stop_flag = False; self.max_mhz = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        return pynvml, h

    def _sample_nvml(self):
        p, h = self.nvml, self.handle
        self.samples.append(float(p.nvmlDeviceGetClockInfo(h, p.NVML_CLOCK_SM)))
        try:
            self.bits |= int(p.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            try:
                self.bits |= int(p.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass

    def _loop_nvml(self):
        while not self.stop_flag:
            try:
                self._sample_nvml()
            except Exception:
                break
            time.sleep(0.005)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self._sample_nvml()                 # probe that the queries work, then drop it: only samples taken under load count
            self.samples.clear(); self.bits = 0
            self.t = threading.Thread(target=self._loop_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def sample_now(self):
        """one synchronous sample from the caller's thread (called while the timed kernels are in flight)"""
        if self.nvml:
            try:
                self._sample_nvml()
            except Exception:
                pass

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            reasons = sorted(name for bit, name in self.REASONS if self.bits & bit)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(self.samples), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


def workload_string(args):
    """config.workload, identical in both arms (the driver compares the strings)"""
    W, H = (args.width or args.frame), (args.height or args.frame)
    if args.workload == "train":
        return (f"train: synthetic {args.volume}^3 volume, example-model.json (8 levels x 8 features, T=2^{args.log2_hashmap}, 64x4 MLP), "
                f"{args.batch} samples per rank per step, fwd + L1 + bwd + Adam")
    return (f"render: synthetic {args.volume}^3 volume, example-model.json (8 levels x 8 features, T=2^{args.log2_hashmap}, 64x4 MLP), "
            f"{W}x{H} frame, macrocell skipping, mode 5 (sample streaming), 16-view orbit")


def build_scene(vnr, dims, train_steps, batch, model_kwargs=None):
    from instantvnr_b200 import synthetic as syn
    gt = syn.make_volume(dims, seed=42)
    vol = vnr.NeuralVolume(vnr.model_json(**(model_kwargs or {})), dims)
    vol.set_groundtruth(gt)
    vol.init_params(1337)
    rgb, alpha = syn.make_tfn(256)
    vol.set_transfer_function(rgb, alpha, (0.0, 1.0))
    # online training with macrocell value ranges learned from the batches (NeuralVolume::train, fast_mode=false)
    done = 0
    while done < train_steps:
        k = min(100, train_steps - done)
        vol.train(k, batch=batch, fast_mode=False)
        done += k
    return vol, gt, (rgb, alpha)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import instantvnr_b200 as vnr
    from instantvnr_b200 import synthetic as syn

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = (args.volume,) * 3
    W, H = (args.width or args.frame), (args.height or args.frame)
    t0 = time.time()
    vol, gt, (rgb, alpha) = build_scene(vnr, dims, args.train_steps, 1 << 16, dict(log2_hashmap=args.log2_hashmap))
    train_step_count, train_loss = vol.stats()
    # N > 1: everything multi-GPU goes through the library's communicator (vnr_comm_init_rank; include/vnr_c.h): attaching the
    # volume replicates rank 0's trained model, macrocell value ranges and sampler stream (independently trained replicas would
    # differ in the last bits: the fp16 reductions of training are order-dependent) and makes vnr_volume_train data parallel;
    # attaching the renderer deals the image strips to the ranks -- finished pixels are stored by the compositing kernels into
    # pinned host frames shared by all ranks (every GPU over its own PCIe link) or, with the download off, into rank 0's device
    # frame over NVLink; a peer-memory barrier kernel closes the frame.  --gather nccl keeps the round-1 harness (torch.distributed
    # gather of padded strips) as the comparison.
    comm = None; tp = None
    use_comm = world > 1 and args.gather != "nccl"
    if use_comm:
        comm = vnr.Comm.init_rank(rank, world, f"vnr-bench-{os.environ.get('MASTER_PORT', '0')}-{os.getppid()}")
        vol.attach_comm(comm)
    elif world > 1:
        from instantvnr_b200.distributed import TileParallelRenderer, broadcast_params
        broadcast_params(vol)
        _, vr, _ = vol.get_macrocell()
        t = torch.from_numpy(vr).cuda(); dist.broadcast(t, src=0); vol.set_macrocell(t.cpu().numpy())

    def make_renderer():
        r = vnr.Renderer(vol)
        r.set_size(W, H)
        r.set_mode(vnr.VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING)
        r.set_sampling_rate(1.0)
        return r

    n_views = 16
    cams = [syn.default_camera(dims, v, n_views) for v in range(n_views)]
    # the frame ring: `--pipeline P` frame slots inside the renderer (own stream, ray / sample buffers and captured wavefront
    # graph each; vnr_renderer_set_frames_in_flight): consecutive vnr_render calls overlap on the device -- the latency-bound tail
    # rounds of frame i run under the head of frame i+1.  Every frame is still rendered completely.
    # frames in flight: --pipeline, default 2 (the reference's double buffer) on one or two GPUs, 3 / 4 on four / eight: a rank's share
    # of the frame is a few short wavefront rounds there and one more frame in flight fills their launch gaps (measured on one
    # GPU rendering 1/8 of the frame: 0.187 / 0.130 / 0.111 / 0.105 ms per frame with 1 / 2 / 3 / 4 in flight; tools/exp_partition_cost.py)
    pipe_default = 2 if world <= 2 else 3 if world <= 4 else 4
    n_pipe = 1 if (world > 1 and not use_comm) else max(1, args.pipeline or pipe_default)
    ren = make_renderer()
    ren.set_frames_in_flight(n_pipe)
    parity = None
    if use_comm:
        # pre-flight parity (the driver's GPU-test box has one GPU): the tile-parallel frame must equal the single-GPU frame bit
        # for bit -- rank 0 renders the whole frame alone first, then the attached renderers render it together
        solo = make_renderer(); solo.set_camera(*cams[1]); solo.render(); want = solo.map_frame(); del solo
        ren.attach_comm(comm)
        ren.set_camera(*cams[1]); ren.render()
        ok = bool(np.array_equal(ren.map_frame(), want)) if rank == 0 else True
        for _ in range(n_pipe - 1):                      # leave every slot of the ring in the same state on every rank
            ren.render()
            if rank == 0:
                ren.map_frame()
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"tile_parallel_frame_equals_single_gpu_frame": bool(flag.item() == 1.0)}
    elif world > 1:
        tp = TileParallelRenderer(ren, mode="nccl")

    def render_frame():
        if tp:
            tp.render()
        else:
            ren.render()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) ----------------
    if tp:
        tp.download = False
    else:
        ren.set_download(False)            # N > 1: pixels now go to rank 0's device frame over NVLink
    pipe_streams = [torch.cuda.ExternalStream(s_) for s_ in ren.streams()]

    def pipelined_frame(i):
        ren.set_camera(*cams[i % n_views])
        render_frame()

    for i in range(max(args.warmup, 2 * n_pipe)):
        pipelined_frame(i)
    barrier()
    clocks = ClockSampler(local); clocks.start()
    decoded = 0; composited = 0; launches = 0; rays = 0
    # The timed region is EXACTLY --steps frames between two barriers + synchronisations.  A short region (the driver's --steps 20
    # is ~12 ms) is repeated -- every repetition is such a region of its own -- and the MEDIAN repetition is reported, so that one
    # slow frame or four clock samples do not decide the line; --steps >= 192 is timed once.
    n_rep = max(1, -(-192 // max(1, args.steps)))
    rep_ms = []
    for _ in range(n_rep):
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in pipe_streams]; ev1 = [torch.cuda.Event(enable_timing=True) for _ in pipe_streams]
        barrier()
        for e, st_ in zip(ev0, pipe_streams):
            e.record(st_)
        for i in range(args.steps):
            pipelined_frame(i)
        for e, st_ in zip(ev1, pipe_streams):
            e.record(st_)
        clocks.sample_now()                             # the queue is still draining: a sample under load even for a very short region
        for st_ in pipe_streams:
            st_.synchronize()
        barrier()
        # all start events were recorded on idle streams at the same moment: the region ends when the last stream finishes
        rep_ms.append(max(ev0[0].elapsed_time(e) for e in ev1))
    ms = float(np.median(rep_ms))
    clk = clocks.stop()
    # samples per frame are deterministic per view: collect the counters outside the timed region, on the
    # same (graph-driven) path that was timed; kernel launches = first round + loop init + 3 per non-empty
    # round + finalize, from the device counters
    prof_steps = min(args.steps, n_views)
    per_view = []
    for i in range(prof_steps):
        ren.set_camera(*cams[i % n_views]); ren.render()
        st = ren.stats(); pr = ren.profile()
        per_view.append((st["samples_decoded"], st["samples_composited"], st["rays_hit"], pr["kernel_launches"]))
    # decode share: the host-enqueued path with CUDA events around every decode launch (same kernels)
    ren.set_profiling(True)
    decode_ms = 0.0; decode_launches = 0; prof_decoded = 0
    for i in range(prof_steps):
        ren.set_camera(*cams[i % n_views]); ren.render()
        st = ren.stats(); pr = ren.profile()
        decode_ms += pr["decode_ms"]; decode_launches += pr["decode_launches"]; prof_decoded += st["samples_decoded"]
    ren.set_profiling(False)
    # frames are deterministic per view: totals of the timed frames follow from the per-view counters
    for i in range(args.steps):
        d_, c_, r_, l_ = per_view[i % n_views % prof_steps]
        decoded += d_; composited += c_; rays += r_; launches += l_
    tot = torch.tensor([float(decoded), float(composited), float(ms), float(decode_ms), float(launches), float(rays)], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); sm = tot.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        decoded, composited, launches, rays = int(sm[0].item()), int(sm[1].item()), int(sm[4].item()), int(sm[5].item())
        ms = mx[2].item()
    value = decoded / (ms * 1e-3)

    # ---------------- end to end through the public call sequence ----------------
    if tp:
        tp.download = True
    else:
        ren.set_download(True)

    def map_frame():
        if tp:
            return tp.map_frame(copy=False)
        return ren.map_frame(copy=False) if rank == 0 else None     # tile-parallel frames are mapped on rank 0

    def e2e_pass():
        for i in range(3):
            ren.set_camera(*cams[i % n_views]); render_frame(); map_frame()
        barrier()
        t_e2e0 = time.perf_counter()
        for i in range(args.steps):
            ren.set_camera(*cams[i % n_views])      # host -> device: the frame constants (kernel arguments)
            render_frame()
            img = map_frame()                       # device -> host: W*H float4 in pinned memory + sync (no extra host copy,
            if img is not None:                     # as vnrRendererMapFrame returns a pointer); touch the result
                checksum = float(img[H // 2, W // 2, 3])
        barrier()                                   # every frame was mapped (synchronised): the host clock brackets the device work
        ms_ = (time.perf_counter() - t_e2e0) * 1e3
        if world > 1:
            t = torch.tensor([ms_], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_ = t.item()
        return ms_

    # the library default: finished pixels are stored straight into the pinned host frame by the compositing kernels
    # (single GPU; the D2H bytes are the same 16 B/pixel, they cross PCIe during the frame instead of after it)
    ms_e2e = float(np.median([e2e_pass() for _ in range(n_rep)]))
    e2e_value = decoded / (ms_e2e * 1e-3)
    ms_e2e_copy = None
    if not tp:
        ren.set_zero_copy(False)                    # comparison: device frame + one cudaMemcpyAsync after the frame (the reference's order)
        ms_e2e_copy = float(np.median([e2e_pass() for _ in range(n_rep)]))
        ren.set_zero_copy(True)
    # the same end-to-end calls with frames in flight (reported next to, not instead of, the strict figure): with a ring of K frame
    # slots inside the renderer, frame i+1 .. i+K-1 are already launched when frame i is mapped (vnr_map_frame returns the oldest
    # unmapped frame); every frame is still mapped (synchronised, host-visible) exactly once
    def set_ring(k):
        """frames in flight of the renderer; a communicator-attached renderer is re-attached (collective, every rank)"""
        if use_comm:
            ren.detach_comm(); ren.set_frames_in_flight(k); ren.attach_comm(comm)
        else:
            ren.set_frames_in_flight(k)

    fps_inflight = {}; inflight = 0
    if not tp:
        K = 3; inflight = K - 1
        set_ring(K)

        def inflight_pass(n):
            for i in range(n + K - 1):
                if i < n:
                    ren.set_camera(*cams[i % n_views]); ren.render()
                if i >= K - 1 and rank == 0:
                    img = ren.map_frame(copy=False)
                    checksum = float(img[H // 2, W // 2, 3])
        for zc in (True, False):
            ren.set_zero_copy(zc)
            inflight_pass(2 * K)
            barrier()
            tq = time.perf_counter()
            inflight_pass(args.steps)
            barrier()
            dt_ = time.perf_counter() - tq
            if world > 1:
                t = torch.tensor([dt_], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt_ = t.item()
            fps_inflight["zero_copy" if zc else "dma_copy"] = args.steps / dt_
        ren.set_zero_copy(True)
        set_ring(n_pipe)

    # ---------------- training throughput of the same volume (steps/s), reported next to the headline ----------------
    # N = 1: vnr_volume_train on the one GPU.  N > 1: the SAME call on every rank of the communicator = synchronous data parallel,
    # 2^18 samples per rank per step (weak scaling), reduce-scatter + Adam + all-gather fused in one kernel over NVLink peer
    # memory; device-timed on every rank between barriers, max over ranks.
    TB = 1 << 18
    vst = torch.cuda.ExternalStream(vol.stream())
    dp_ok = None
    if world == 1 or use_comm:
        vol.train(5, batch=TB, fast_mode=True)
        barrier()
        et0, et1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        et0.record(vst)
        vol.train(20, batch=TB, fast_mode=True)
        et1.record(vst)
        vst.synchronize(); barrier()
        train_ms = et0.elapsed_time(et1) / 20
        if world > 1:
            t = torch.tensor([train_ms], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); train_ms = t.item()
            # pre-flight parity of the data-parallel path: replicas bit-identical after the steps above, the same global loss on every rank
            p16 = torch.from_numpy(vol.get_params_f16().astype(np.int64)).cuda()
            sig = torch.stack([p16.sum(), (p16 * (torch.arange(p16.numel(), device="cuda") % 8191 + 1)).sum()]).double()
            lo_, hi_ = sig.clone(), sig.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            loss_t = torch.tensor([vol.last_loss()], device="cuda", dtype=torch.float64); l0, l1 = loss_t.clone(), loss_t.clone()
            dist.all_reduce(l0, op=dist.ReduceOp.MIN); dist.all_reduce(l1, op=dist.ReduceOp.MAX)
            dp_ok = bool(torch.equal(lo_, hi_)) and bool(l0.item() == l1.item()) and bool(np.isfinite(l0.item()))
            parity["data_parallel_replicas_bit_identical"] = dp_ok
    else:
        train_ms = float("nan")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = measured_peaks()
    # ---- roofline of the dominant kernel on this rank (decode): event-timed launches.
    # The kernel is a gather: 64 scattered 16-byte loads per sample out of the hash table.  What bounds it is the rate at which an
    # SM's L1TEX turns scattered requests into L2 sector fetches -- ~1 per cycle per SM (csrc/probe.cu: plain ld.global.nc.v4 at
    # random offsets, nothing of the product's gather code; it does NOT depend on the thread count, 256 threads per SM reach it,
    # but it halves when the CTA's shared memory selects certain shared-memory / L1 splits: tools/exp_probe_cta.py, DESIGN 3.1,
    # which is why the decode ring holds 7 tiles and the training kernel stays under 192 KB).  l2_gather_gbs is that probe over a
    # flat 46.7 MB buffer, hbm_gather_gbs the same over 307 MB (every load a DRAM miss).
    # `achieved` / `frac` are like for like with the probe: the SAME kernel on uniform random coordinates (launches timed here
    # with CUDA events), 1024 algorithmic gather bytes per sample.  The launches INSIDE the frames run faster
    # still (`in_frame`): neighbouring rows are the same step of neighbouring rays, lanes of a warp ask for the same table
    # entries and those requests collapse before the L1TEX -- their algorithmic bytes overcount what the bound unit sees, so
    # they are reported next to the roofline, not as its fraction.
    GATHER_BYTES = 1024
    table_bytes = (vol.n_params - vol.n_mlp_params) * 2
    probe_ops = (1 << 22) * 64
    l2_ms, _ = vnr.probe_memory("loads", 2920448 * 16, probe_ops, 3)
    hbm_ms, _ = vnr.probe_memory("loads", 19173376 * 16, probe_ops, 3)
    own_ms, _ = vnr.probe_memory("loads", table_bytes, probe_ops, 3)
    gbs = lambda ms_: probe_ops * 16 / (ms_ * 1e-3) / 1e9
    gather_peak = gbs(own_ms)
    # the same loads with the decode's level structure (8 per level, uniformly inside that level's table): coarse levels stay
    # cache-resident, levels larger than the L2 miss in proportion -- the like-for-like ceiling when the table exceeds the L2
    lv_entries = vnr.hash_grid_level_entries(log2_hashmap=args.log2_hashmap)
    lv_ms, _ = vnr.probe_levels(lv_entries, 1 << 22, 3)
    levels_gbs = (1 << 22) * 64 * 16 / (lv_ms * 1e-3) / 1e9
    # `peak`: the level-structured probe while the table is L2-resident (like for like with the kernel on uniform coordinates: the
    # coarse levels hit in L1 in both).  A table beyond the L2 (T = 2^22): the request-rate ceiling of an L2-resident table stays
    # the upper bound and `frac` says how far DRAM latency on the missing levels keeps the kernel below it; the two DRAM-side probes
    # (levels_gather_gbs, hbm_gather_gbs) are reported next to it -- the kernel exceeds both, its x-neighbour corners share sectors
    gather_peak = levels_gbs if table_bytes <= 100e6 else gbs(l2_ms)
    bound = "l1tex_gather"
    decode_rate = prof_decoded / (decode_ms * 1e-3) if decode_ms > 0 else 0.0
    achieved = decode_rate * GATHER_BYTES / 1e9
    # the same kernel on uniform random coordinates (no coherence between neighbouring rows): apples to apples with the probe
    nu = 1 << 22
    xyz_u = torch.rand(nu, 3, device="cuda"); out_u = torch.empty(nu, device="cuda")
    vst = torch.cuda.ExternalStream(vol.stream())
    for _ in range(2):
        vol.decode(xyz_u, out_u, nu)
    eu0, eu1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    eu0.record(vst)
    for _ in range(5):
        vol.decode(xyz_u, out_u, nu)
    eu1.record(vst); vst.synchronize()
    uniform_rate = 5 * nu / (eu0.elapsed_time(eu1) * 1e-3)
    # DRAM bytes of the largest decode launch of a frame from the committed `ncu --set full` capture of THIS configuration (per
    # launch), next to the algorithmic bytes of that same launch; null when no capture of this configuration is committed
    traffic = None; traffic_detail = None
    tpath = os.path.join(ROOT, "profiles", f"decode_traffic_t{args.log2_hashmap}_{W}x{H}.json")
    if world == 1 and os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes"]
            traffic_detail = {"unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "algorithmic_bytes_of_that_launch": tj["samples"] * BYTES_PER_SAMPLE,
                              "samples_of_that_launch": tj["samples"], "launch_us_under_ncu": tj["duration_us"], "source": tj["source"]}
        except Exception:
            pass
    achieved_u = uniform_rate * GATHER_BYTES / 1e9
    roofline = {"bound": bound, "achieved": round(achieved_u, 1), "peak": round(gather_peak, 1), "unit": "GB/s", "frac": round(achieved_u / gather_peak, 4),
                "traffic": traffic, "traffic_detail": traffic_detail,
                "kernel": "decode_kernel<8,.> (fused hash-grid gather + tcgen05 MLP)",
                "achieved_source": "decode_kernel on 2^22 uniform random coordinates, 5 launches timed with CUDA events in this run; 1024 algorithmic gather bytes per sample",
                "peak_source": ("measured in this run: random 16-byte ld.global.nc, 8 per level inside each level's own table (csrc/probe.cu vnr_probe_levels), full occupancy, "
                                "nothing of the product's gather code" if table_bytes <= 100e6 else
                                "measured in this run: random 16-byte ld.global.nc over an L2-resident 46.7 MB buffer (the request-rate ceiling); this run's table exceeds the L2"),
                "table_exceeds_l2": bool(table_bytes > 100e6),
                "l2_gather_gbs": round(gbs(l2_ms), 1), "hbm_gather_gbs": round(gbs(hbm_ms), 1), "levels_gather_gbs": round(levels_gbs, 1), "table_bytes": table_bytes,
                "algorithmic_bytes_per_sample": GATHER_BYTES, "decode_uniform_samples_per_sec": uniform_rate,
                "in_frame": {"decode_samples_per_sec": decode_rate, "gather_gbs": round(achieved, 1), "ratio_to_peak": round(achieved / gather_peak, 4),
                             "decode_ms_per_frame": round(decode_ms / prof_steps, 4), "decode_launches_per_frame": decode_launches / prof_steps,
                             "note": "the decode launches of the frames (event-timed, host-enqueued rounds of the profiling pass): coherent coordinates, requests of a warp "
                                     "collapse before the L1TEX, so the algorithmic bytes exceed what the bound unit serves"},
                "decode_samples_per_sec": decode_rate, "decode_ms_per_frame": round(decode_ms / prof_steps, 4),
                "hbm_copy_peak": peak, "hbm_copy_peak_source": peak_kind,
                "hbm_copy_frac": round(decode_rate * BYTES_PER_SAMPLE / 1e9 / peak, 4),
                "note": "hbm_copy_frac (all 1044 algorithmic B/sample of the in-frame launches against the HBM copy peak) is kept for comparison with round 1; the table is "
                        "read from L2, not HBM, when it fits (traffic = DRAM bytes of one launch from the committed ncu capture of this configuration)"}

    # the CPU baseline is measured at N=1 only (under torchrun the host cores are shared by the ranks)
    cpu = cpu_baseline(vol, dims, cams, rgb, alpha, args) if world == 1 else {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "measured at N=1 only"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": workload_string(args),
                   "weights": f"trained here for {train_step_count} steps (batch 2^16), mean L1 loss {train_loss:.4f}",
                   "l2_flush": "inputs larger than L2: per-frame sample/value/ray-state buffers (~500 MB) stream through the 126 MB L2 between frames",
                   "parallelism": f"tile-parallel x{world}" if world > 1 else "single GPU",
                   "timed_region": f"exactly {args.steps} frames between barriers + synchronisations" + (f", repeated {n_rep} times, median repetition reported" if n_rep > 1 else ""),
                   "frames_in_flight": f"{n_pipe} frame slot(s) inside the renderer for the device-resident `value` (vnr_renderer_set_frames_in_flight; the strict end-to-end pass maps every frame before the next is launched)"},
        "fps": args.steps / (ms * 1e-3), "samples_per_frame": decoded / args.steps, "composited_per_frame": composited / args.steps,
        "rays_hit_per_frame": rays / args.steps,
        "train_steps_per_sec_batch_2p18": 1000.0 / train_ms, "train_samples_per_sec": world * (1 << 18) * 1000.0 / train_ms,
        "dp_steps_per_sec": (1000.0 / train_ms) if world > 1 else None, "dp_global_batch": world * (1 << 18),
        "dp_mode": ("sharded optimizer over NVLink peer memory (reduce-scatter + Adam + all-gather in one kernel) behind vnr_volume_train on a communicator; "
                    "2^18 samples per rank per step (weak scaling)") if world > 1 else "single GPU",
        "parity_checked": (bool(parity) and all(parity.values())) if world > 1 else None, "parity": parity,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 480, "d2h_bytes_per_step": W * H * 16, "fps": args.steps / (ms_e2e * 1e-3),
                "d2h_note": f"the mapped host frame is {W}x{H} float4 = {W * H * 16} bytes; with the zero-copy download a pixel that is zero and was zero in that host "
                            f"buffer is not stored again, so about 16 B x {int(rays / max(1, args.steps))} (rays that hit the volume) cross PCIe per frame on this orbit; "
                            "fps_copy_after_frame moves all of it with one cudaMemcpyAsync",
                "frame_path": ("tile-parallel NCCL gather + download on rank 0" if tp else
                               "tile-parallel: every rank's compositing kernels store its strips into ONE pinned host frame shared by all ranks (each GPU over its own PCIe link); rank 0 maps it"
                               if use_comm else "zero-copy: compositing kernels store finished pixels into the pinned host frame"),
                "fps_copy_after_frame": (args.steps / (ms_e2e_copy * 1e-3)) if ms_e2e_copy else None,
                "fps_with_frames_in_flight": (max(fps_inflight.values()) if fps_inflight else None), "fps_with_frames_in_flight_by_download": fps_inflight,
                "frames_in_flight": inflight,
                "note": "value / fps: strict loop, every frame mapped before the next one is launched.  fps_with_frames_in_flight: the same two calls on ONE renderer whose "
                        "frame ring holds 3 slots (vnr_renderer_set_frames_in_flight): vnr_map_frame returns the oldest unmapped frame while the next two compute"},
        "gpu_launches": int(launches),
        "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "setup_seconds": round(time.time() - t0, 1),
    }
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(vol, dims, cams, rgb, alpha, args):
    """The CPU oracle (scalar port, OpenMP over rays) on a bounded sample of the same workload: whole frames
    of the same scene, view after view, until about `--cpu-seconds` of CPU work."""
    import oracle as O
    O.use_host_cores()
    m = O.ModelCfg()
    p16 = vol.get_params_f16()
    md, vr, mo = vol.get_macrocell()
    w, h = (args.width or args.frame), (args.height or args.frame)
    m = O.ModelCfg(log2_hashmap=args.log2_hashmap)
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    t0 = time.perf_counter()
    n = 0; views = 0
    while views < len(cams) and time.perf_counter() - t0 < args.cpu_seconds:
        fr = O.Frame(dims, w, h, *cams[views])
        _, _, st = O.render(m, p16, fr, mo, colors, alpha, acc_mode=0)
        n += st["samples_decoded"]; views += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": O.lib().orc_num_threads(), "kind": "port",
            "sample": f"{views} frame(s) {w}x{h} of the same scene ({n} samples, {dt:.1f} s); fps {views / dt:.3f}"}


def synth_volume_device(dims, seed=42):
    """the synthetic volume of instantvnr_b200/synthetic.py, evaluated on the device slab by slab (torch is plumbing
    here: a 1024^3 volume is 4 GiB and lives in HBM only)"""
    import torch
    dx, dy, dz = dims
    rng = np.random.RandomState(seed)
    centres = rng.uniform(0.2, 0.8, size=(8, 3)).astype(np.float32)
    sigmas = rng.uniform(0.05, 0.15, size=8).astype(np.float32)
    amps = rng.uniform(0.5, 1.0, size=8).astype(np.float32)
    x = ((torch.arange(dx, device="cuda", dtype=torch.float32) + 0.5) / dx)[None, None, :]
    y = ((torch.arange(dy, device="cuda", dtype=torch.float32) + 0.5) / dy)[None, :, None]
    vol = torch.empty(dz, dy, dx, device="cuda", dtype=torch.float32)
    slab = 16
    for k0 in range(0, dz, slab):
        z = ((torch.arange(k0, min(dz, k0 + slab), device="cuda", dtype=torch.float32) + 0.5) / dz)[:, None, None]
        s = 0.05 * (torch.sin(2 * np.pi * x) * torch.sin(2 * np.pi * y) * torch.sin(2 * np.pi * z) + 1.0)
        for c, sg, a in zip(centres, sigmas, amps):
            s = s + float(a) * torch.exp(-((x - float(c[0])) ** 2 + (y - float(c[1])) ** 2 + (z - float(c[2])) ** 2) / float(2 * sg * sg))
        vol[k0:k0 + slab] = s
    lo, hi = vol.min(), vol.max()
    vol.sub_(lo).div_(hi - lo)
    return vol


def run_train(args):
    """BASELINE configs[2] (1 GPU) / configs[3] (N GPUs): online training steps/s -- per step and per rank: draw
    --batch samples of the ground truth, forward + L1 + backward, (N > 1: all-reduce of hash-grid + MLP gradients),
    Adam.  The volume lives in HBM (1024^3 float = 4 GiB).  value = optimizer steps per second; weak scaling (the
    per-rank batch is fixed, the global batch grows with N)."""
    import torch
    import torch.distributed as dist
    import instantvnr_b200 as vnr
    from instantvnr_b200.distributed import DataParallelTrainer, GpuTrainBackend, broadcast_params

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = (args.volume,) * 3
    vol = vnr.NeuralVolume(vnr.model_json(log2_hashmap=args.log2_hashmap), dims)
    ooc_info = None
    if args.out_of_core:
        # BASELINE configs[3]: the volume stays in a raw file (uint8, written once per box from the same procedural volume);
        # every rank keeps its own pool of random slabs of it in HBM, refreshed by num_concurrent_blocks slabs per step -- pulled by
        # the GPU itself out of the page cache (the file mapped and registered; csrc/slab_sampler.cu) -- and samples the pool on
        # the device (OutOfCoreSampler, neural_sampler.cpp:1065-1120)
        path = os.path.join(os.environ.get("VNR_BENCH_TMP", "/tmp"), f"vnr_bench_volume_{args.volume}_u8.raw")
        if local == 0 and not (os.path.exists(path) and os.path.getsize(path) == args.volume ** 3):
            gt = synth_volume_device(dims)
            with open(path + ".tmp", "wb") as f:
                for z0 in range(0, dims[2], 64):
                    f.write((gt[z0:z0 + 64] * 255.0 + 0.5).clamp_(0, 255).to(torch.uint8).cpu().numpy().tobytes())
            os.replace(path + ".tmp", path)
            del gt
        if world > 1:
            dist.barrier()
        vol.set_groundtruth_outofcore(path, "uint8", (0.0, 255.0))
        ooc_info = vol.outofcore_info()
    else:
        gt = synth_volume_device(dims)
        vol.set_groundtruth_device(gt)
        del gt
    torch.cuda.empty_cache()
    vol.init_params(1337)
    # the public call (vnrNeuralVolumeTrain -> vnr_volume_train) is what is timed.  N > 1, --dp-mode sharded: the same call on every
    # rank of the library's communicator (synchronous data parallel; reduce-scatter + Adam + all-gather in one kernel over NVLink
    # peer memory).  --dp-mode allreduce keeps the torch.distributed harness (NCCL all-reduce of the gradient buffers) as comparison.
    comm = None; dp = None
    if world > 1 and args.dp_mode == "sharded":
        comm = vnr.Comm.init_rank(rank, world, f"vnr-bench-train-{os.environ.get('MASTER_PORT', '0')}-{os.getppid()}")
        vol.attach_comm(comm)
    elif world > 1:
        broadcast_params(vol)
        dp = DataParallelTrainer(GpuTrainBackend(vol), mode=args.dp_mode)
    stream = torch.cuda.ExternalStream(vol.stream())
    n = args.batch

    def steps_(k, want_loss=False):
        if dp is not None:
            loss = None
            for _ in range(k):
                loss = dp.step(n, want_loss=want_loss)
            return loss
        vol.train(k, batch=n, fast_mode=True)
        return vol.last_loss() if want_loss else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps_(max(args.warmup, 3))
    barrier()
    clocks = ClockSampler(local); clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    steps_(args.steps)
    ev1.record(stream)
    clocks.sample_now()                                 # the queue is still draining: a sample under load even for a very short region
    stream.synchronize(); barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    # end to end: one step per call, with the per-step result (the loss) read back to the host before the next call
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = steps_(1, want_loss=True)
    stream.synchronize(); barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t[0].item(), t[1].item()
    step_count, mean_loss = vol.stats()
    psnr = vol.psnr() if args.volume <= 512 and not args.out_of_core else None
    if ooc_info is not None:
        up0 = ooc_info["bytes_uploaded"]; ooc_info = vol.outofcore_info(); ooc_info["bytes_uploaded_per_step"] = (ooc_info["bytes_uploaded"] - up0) / max(1, step_count)
    # the fused kernel against ITS bound: the SM's L1TEX request stage serves the gather's scattered loads and the backward's
    # scattered fp16x8 reductions one after the other, so the floor of the kernel is what an independent probe needs for the
    # same number of random 16-byte loads plus random 16-byte reductions over tables of the same size (csrc/probe.cu kind
    # "mixed": n x 64 of each, full occupancy, no MLP, no tiles) -- measured here, next to the kernel on a fixed batch
    kernel_us = floor_us = None
    if world == 1 and rank == 0 and not args.out_of_core:
        xyz = torch.empty(n, 3, device="cuda"); tgt = torch.empty(n, device="cuda")
        vol.sample(xyz, tgt, n); torch.cuda.synchronize()
        for _ in range(3):
            vol.train_grads(xyz, tgt, n, n)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for _ in range(20):
            vol.train_grads(xyz, tgt, n, n)
        k1.record(stream); stream.synchronize()
        vol.optimizer_step(); torch.cuda.synchronize()
        kernel_us = k0.elapsed_time(k1) / 20 * 1e3
        floor_ms, _ = vnr.probe_memory("mixed", (vol.n_params - vol.n_mlp_params) * 2, n * 64, 5)
        floor_us = floor_ms * 1e3
    if rank == 0:
        peak, peak_kind = measured_peaks()
        n_grid = vol.n_params - vol.n_mlp_params
        # algorithmic bytes per step and rank: per sample 1024 B gather + 1024 B gradient reduction + 16 B sample;
        # optimizer sweep 38 B per touched parameter (upper bound: all) + gradient clear
        bytes_step = n * (1024 + 1024 + 16) + n_grid * 38
        achieved = bytes_step / (ms / args.steps * 1e-3) / 1e9
        out = {"metric": "train_steps_per_sec", "value": args.steps / (ms * 1e-3), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f16", "data": "synthetic",
               "config": {"workload": workload_string(args),
                          "path": ("out-of-core: raw uint8 file on local disk, a per-rank pool of random slabs in HBM, 1024 slabs (107 MB) per rank and step pulled by the GPU out of the page cache over PCIe, sampled on the device"
                                   if args.out_of_core else "volume resident in HBM, sampled on the device") + ("" if world == 1 else ", gradient all-reduce (fp16 grid + fp32 MLP) over NCCL + replicated Adam" if dp is not None
                                                               else ", optimizer fused with its collectives over NVLink peer memory (reduce-scatter + Adam + all-gather in one kernel)"),
                          "global_batch": n * world, "l2_flush": "per-step parameter-state sweep (~0.9 GB) exceeds L2",
                          "parallelism": f"dp{world}"},
               "out_of_core": ooc_info,
               "samples_per_sec": args.steps * n * world / (ms * 1e-3), "mean_loss": mean_loss, "last_loss": loss, "volume_psnr_db": psnr,
               "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                       "note": "samples are drawn on the device from the HBM-resident volume (the reference's StaticSampler does the same); the loss is read back every step"},
               "gpu_launches": args.steps * 9, "clocks": clk,
               "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None,
                            "kernel": "whole step (train_step_kernel + adam_grid_kernel)", "peak_source": peak_kind, "algorithmic_bytes_per_step": bytes_step,
                            "train_step_kernel": None if kernel_us is None else {
                                "bound": "l1tex_gather + l1tex_reductions", "kernel_us": round(kernel_us, 1), "floor_us": round(floor_us, 1), "frac": round(floor_us / kernel_us, 4),
                                "note": "kernel_us: train_grads (fused forward + loss + backward kernel and the reduction of its per-CTA weight-gradient partials) on a fixed batch, "
                                        "20 calls timed with CUDA events; floor_us: the probe's n x 64 random 16-byte loads + n x 64 random fp16x8 reductions in one launch"}}}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_interleaved(args):
    """BASELINE configs[2] as the interactive app runs it (apps/int_dual_volume.cpp:636-671): every iteration is one online
    training step (2^18 samples, macrocell value ranges updated from the batch, max opacity refreshed) followed by one
    1024^2 frame of the volume as trained so far, mapped to the host.  value = iterations (= frames = steps) per second."""
    import torch
    import instantvnr_b200 as vnr
    from instantvnr_b200 import synthetic as syn
    torch.cuda.set_device(0)
    dims = (args.volume,) * 3
    W, H = (args.width or args.frame), (args.height or args.frame)
    vol, gt, (rgb, alpha) = build_scene(vnr, dims, 50, 1 << 16, dict(log2_hashmap=args.log2_hashmap))
    ren = vnr.Renderer(vol)
    ren.set_size(W, H)
    ren.set_mode(vnr.VNR_RAYMARCHING_NO_SHADING_SAMPLE_STREAMING)
    cams = [syn.default_camera(dims, v, 16) for v in range(16)]
    n = args.batch

    def iteration(i):
        vol.train(1, batch=n, fast_mode=False)
        ren.set_camera(*cams[i % 16])
        ren.render()
        return ren.map_frame(copy=False)

    for i in range(max(args.warmup, 3)):
        iteration(i)
    torch.cuda.synchronize()
    clocks = ClockSampler(0); clocks.start()
    decoded = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        img = iteration(i)
        decoded += ren.stats()["samples_decoded"]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    clk = clocks.stop()
    step_count, mean_loss = vol.stats()
    emit({"metric": "interleaved_train_render_iterations_per_sec", "value": args.steps / dt, "unit": "iterations/s", "n_gpus": 1, "steps": args.steps,
          "warmup": max(args.warmup, 3), "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f16", "data": "synthetic",
          "config": {"workload": f"interleaved: per iteration one training step ({n} samples, fwd + L1 + bwd + Adam + macrocell update) and one {W}x{H} "
                                 f"mode-5 frame of the synthetic {args.volume}^3 volume, frame mapped to the host (host-timed, every frame synchronised)",
                     "parallelism": "single GPU"},
          "train_samples_per_sec": args.steps * n / dt, "render_samples_per_sec": decoded / dt, "training_step": step_count, "mean_loss": mean_loss,
          "volume_psnr_db": vol.psnr(), "e2e": {"value": args.steps / dt, "unit": "iterations/s", "h2d_bytes_per_step": 480, "d2h_bytes_per_step": W * H * 16},
          "gpu_launches": None, "clocks": clk})


def run_reference_train(args):
    """reference arm of the train workload: the reference's own Trainer::training_step (tiny-cuda-nn built unmodified from
    /root/reference/tcnn, oracle/_ref) on the same GPU, on batches pre-drawn with torch (the reference's StaticSampler needs
    the OVR framework).  Per step: the reference's training step alone; no kernel of this repo runs in this arm."""
    import torch
    import instantvnr_b200 as vnr
    from oracle import tcnn_ref
    base = {"metric": "train_steps_per_sec", "unit": "steps/s", "n_gpus": args.gpus, "gpus_used": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic", "impl": "reference"}
    if not (torch.cuda.is_available() and tcnn_ref.available()):
        base.update({"unavailable": "the reference tcnn build (oracle/_ref) or a GPU is missing; the training step has no CPU implementation in the reference"})
        emit(base); return
    dims = (args.volume,) * 3
    gt = synth_volume_device(dims)
    ref = tcnn_ref.RefNetwork(vnr.model_json(log2_hashmap=args.log2_hashmap), 1337)
    n = args.batch
    st = torch.cuda.Stream()
    # a ring of pre-drawn batches (uniform coordinates, trilinear targets: what StaticSampler's generate_random + tex3D produce).
    # The reference's own sampler needs the OVR framework; drawing outside the timed region charges the reference nothing for it.
    g = gt.view(1, 1, dims[2], dims[1], dims[0])
    ring = []
    for _ in range(8):
        xyz = torch.rand(n, 3, device="cuda")
        tgt = torch.nn.functional.grid_sample(g, (xyz * 2 - 1).view(1, 1, 1, n, 3), mode="bilinear", padding_mode="border", align_corners=False).view(n).contiguous()
        ring.append((xyz, tgt))
    del g, gt
    torch.cuda.synchronize()
    counter = [0]

    def step(want_loss=False):
        xyz, tgt = ring[counter[0] % len(ring)]; counter[0] += 1
        return ref.training_step(xyz, tgt, n, st.cuda_stream, want_loss=want_loss)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st); st.synchronize()
    ms = e0.elapsed_time(e1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(want_loss=True)
    st.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    v = args.steps / (ms * 1e-3)
    base.update({"value": v, "ms_per_step": ms / args.steps, "last_loss": loss,
                 "config": {"workload": workload_string(args),
                            "path": "the reference's Trainer::training_step (tcnn, CUDA-graph captured fwd+loss+bwd, Adam) on the same B200; batches pre-drawn (sampling not timed)",
                            "global_batch": n, "parallelism": "dp1"},
                 "cpu_baseline": {"value": v, "unit": "steps/s", "cores": 0, "kind": "reference", "sample": f"{args.steps} steps of {n} samples on the same GPU"},
                 "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4}})
    emit(base)


def reference_round_coords(dims, W, H, view=1, n_views=16, n_iters=16, fovy=60.0):
    """Coordinates of the first wavefront round as the reference's marcher writes them (renderer.cpp:87-96 camera basis,
    method_raymarching.cu:658-730): for every camera ray that hits the volume box, `n_iters` unit steps from the entry point,
    stored step-major over the live rays (coords[numRays * k + ray]).  Pure torch; none of this repo's kernels."""
    import torch
    from instantvnr_b200 import synthetic as syn
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    cam_from, cam_at, cam_up = [torch.tensor(v, device=dev, dtype=torch.float32) for v in syn.default_camera(dims, view, n_views)]
    d = torch.tensor(dims, device=dev, dtype=torch.float32)
    t = 2.0 * float(np.tan(fovy * 0.5 * np.pi / 180.0)); aspect = W / float(H)
    fwd = torch.nn.functional.normalize(cam_at - cam_from, dim=0)
    hor = t * aspect * torch.nn.functional.normalize(torch.linalg.cross(fwd, cam_up), dim=0)
    ver = torch.linalg.cross(hor, fwd) / aspect
    ix = (torch.arange(W, device=dev, dtype=torch.float32) + 0.5) / W - 0.5
    iy = (torch.arange(H, device=dev, dtype=torch.float32) + 0.5) / H - 0.5
    dirs = torch.nn.functional.normalize(fwd[None, None, :] + ix[None, :, None] * hor[None, None, :] + iy[:, None, None] * ver[None, None, :], dim=-1).reshape(-1, 3)
    inv = 1.0 / dirs
    lo = (-d / 2 - cam_from) * inv; hi = (d / 2 - cam_from) * inv
    t0 = torch.minimum(lo, hi).amax(-1).clamp_min(0.0); t1 = torch.maximum(lo, hi).amin(-1)
    live = t1 > t0
    dirs, t0, t1 = dirs[live], t0[live], t1[live]
    k = torch.arange(n_iters, device=dev, dtype=torch.float32)[:, None]
    tt = torch.minimum(t0[None, :] + k + 0.5, t1[None, :])                       # [n_iters][live rays]: step-major
    p = cam_from[None, None, :] + tt[:, :, None] * dirs[None, :, :]
    return ((p + d / 2) / d).clamp(0.0, 1.0).reshape(-1, 3).contiguous()


def run_reference_frames(args, base):
    """The reference's OWN frame pipeline on this GPU: its ray marcher (core/renderer/method_raymarching.cu, mode 5: raygen ->
    [intersect -> NeuralVolume::inference -> compose] with a host sync per round), its macrocells (core/macrocell.cu) and its
    tiny-cuda-nn decode and training step -- all compiled unmodified from /root/reference (oracle/ref_marcher, oracle/ref_driver;
    only the un-vendored OVR headers are stood in for).  Same scene as our arm: the synthetic volume, 600 training steps of
    batch 2^16 (here through the reference's Trainer on torch-drawn samples), the 256-entry transfer function, the 16-view orbit,
    every frame downloaded to the host (refm_render copies the frame and synchronises, as vnrRender + vnrRendererMapFrame).
    value = network evaluations per second as the reference issues them (16 slots per live ray and round, padded to 256)."""
    import torch
    import instantvnr_b200 as vnr            # model_json / synthetic helpers only: no kernel of this repo runs in this arm
    from instantvnr_b200 import synthetic as syn
    from oracle import marcher_ref, tcnn_ref
    dims = (args.volume,) * 3
    W, H = (args.width or args.frame), (args.height or args.frame)
    gt = syn.make_volume(dims, seed=42)
    net = tcnn_ref.RefNetwork(vnr.model_json(log2_hashmap=args.log2_hashmap), 1337)
    # training samples: uniform coordinates, trilinear targets of the normalised volume (what StaticSampler's tex3D returns)
    g = torch.from_numpy(gt).cuda().view(1, 1, dims[2], dims[1], dims[0])
    nb = 1 << 16
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(args.train_steps):
            xyz = torch.rand(nb, 3, device="cuda")
            tgt = torch.nn.functional.grid_sample(g, (xyz * 2 - 1).view(1, 1, 1, nb, 3), mode="bilinear", padding_mode="border", align_corners=False).view(nb).contiguous()
            loss = net.training_step(xyz, tgt, nb, st.cuda_stream, want_loss=False)
    st.synchronize()
    ref = marcher_ref.RefMarcher(dims, gt)
    rgb, alpha = syn.make_tfn(256)
    ref.set_transfer_function(rgb, alpha, (0.0, 1.0))
    ref.set_decoder(marcher_ref.function_address(tcnn_ref.lib(), "ref_inference"), net.h)
    cams = [syn.default_camera(dims, v, 16) for v in range(16)]
    for i in range(max(3, args.warmup)):
        ref.reset_accumulation(); ref.render(5, (W, H), *cams[i % 16], neural=True)
    torch.cuda.synchronize()
    coords = 0; calls = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        ref.reset_accumulation()
        img, rs = ref.render(5, (W, H), *cams[i % 16], neural=True)
        coords += rs["decode_coords"]; calls += rs["decode_calls"]
    dt = time.perf_counter() - t0
    # USEFUL samples per frame (what our arm's `value` counts: samples that are actually composited / decoded for a ray), from the
    # CPU oracle marching the same views with the same weights, macrocells and transfer function -- the reference's marcher itself
    # only reports how many coordinates it pushed through the network (16 slots per live ray and round, used or not).
    import oracle as O
    O.use_host_cores()                            # torchrun exports OMP_NUM_THREADS=1
    m = O.ModelCfg(8, 8, args.log2_hashmap, 16, 2.0, 4)
    p16 = net.get_params_f16()
    _, _, mo = ref.get_macrocell()
    colors = np.concatenate([rgb, np.ones((rgb.shape[0], 1), np.float32)], 1)
    n_count = min(16, args.steps)
    useful_per_view = [None] * n_count
    stride = 1
    for vw in range(n_count):
        if vw % stride:
            continue
        tc = time.perf_counter()
        fr = O.Frame(dims, W, H, *cams[vw])
        _, _, ost = O.render(m, p16, fr, mo, colors, alpha, acc_mode=1)
        useful_per_view[vw] = ost["samples_decoded"]
        if vw == 0 and time.perf_counter() - tc > 4.0:
            stride = 4                            # a slow host: count every fourth view, the others take their neighbour's count
    last = useful_per_view[0]
    for vw in range(n_count):
        if useful_per_view[vw] is None:
            useful_per_view[vw] = last
        last = useful_per_view[vw]
    useful = sum(useful_per_view[i % n_count] for i in range(args.steps))
    v = useful / dt
    base.update({"value": v, "ms_per_step": dt * 1e3 / args.steps, "fps": args.steps / dt, "scaling": "strong",
                 "config": {"workload": workload_string(args),
                            "path": "the reference's own marcher + macrocell + tiny-cuda-nn sources (compiled unmodified from /root/reference), every frame downloaded",
                            "weights": f"trained here for {args.train_steps} steps (batch 2^16) by the reference's Trainer::training_step"},
                 "samples_per_frame": useful / args.steps, "network_evaluations_per_frame": coords / args.steps, "network_evaluations_per_sec": coords / dt,
                 "wavefront_rounds_per_frame": calls / args.steps,
                 "cpu_baseline": {"value": v, "unit": UNIT, "cores": 0, "kind": "reference",
                                  "sample": f"{args.steps} frames of the same workload on the same B200 (the reference has no CPU path)"},
                 "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": W * H * 16, "fps": args.steps / dt},
                 "note": "value = useful samples per second (samples a ray actually takes; counted by the CPU oracle on the same views and weights, outside the "
                         "timed region), the quantity our arm reports; the reference pushes 16 slots per live ray and round through its network, used or "
                         "not: network_evaluations_per_sec"})
    emit(base)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "train":
        return run_reference_train(args)
    import oracle as O
    from oracle import tcnn_ref
    n = 1 << 24
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "gpus_used": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "impl": "reference"}
    try:
        import torch
        have_gpu = torch.cuda.is_available() and tcnn_ref.available()
    except Exception:
        have_gpu = False
    try:
        from oracle import marcher_ref
        have_marcher = have_gpu and marcher_ref.available()
    except Exception:
        have_marcher = False
    if have_marcher:
        return run_reference_frames(args, base)
    if have_gpu:
        import instantvnr_b200 as vnr
        ref = tcnn_ref.RefNetwork(vnr.example_model_json(), 1337)
        st = torch.cuda.Stream()
        torch.manual_seed(0)

        def time_decode(xyz):
            useful = xyz.shape[0]
            pad = (-useful) % 256                      # NeuralVolume::inference pads the batch to a multiple of 256 (network.cu:1043-1052)
            if pad:
                xyz = torch.cat([xyz, xyz[-1:].expand(pad, 3)]).contiguous()
            cnt = xyz.shape[0]
            out = torch.empty(cnt, device="cuda")
            with torch.cuda.stream(st):
                for _ in range(max(3, args.warmup)):
                    ref.inference(xyz, out, cnt, st.cuda_stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(args.steps):
                    ref.inference(xyz, out, cnt, st.cuda_stream)
                e1.record(st)
            st.synchronize()
            ms_ = e0.elapsed_time(e1)
            return useful * args.steps / (ms_ * 1e-3), ms_ / args.steps

        # (a) the decode batch of a wavefront round laid out as the reference's marcher lays it out (iterative_intersect_kernel,
        # method_raymarching.cu:687-730: coords[numRays * k + ray], k < 16): camera rays of the same orbit, 16 unit steps from the
        # entry point, live rays only.  Built with torch ops -- neighbouring rows are neighbouring rays at the same step, the
        # coherence the reference's own decode sees inside a frame.
        dims = (args.volume,) * 3
        W, H = (args.width or args.frame), (args.height or args.frame)
        xyz_frame = reference_round_coords(dims, W, H, view=1)
        v_frame, ms_frame = time_decode(xyz_frame)
        # (b) uniform random coordinates (no coherence): the lower bracket
        n = 1 << 24
        v_uniform, ms_uniform = time_decode(torch.rand(n, 3, device="cuda"))
        v, ms_step = (v_frame, ms_frame) if v_frame >= v_uniform else (v_uniform, ms_uniform)
        base.update({"value": v, "ms_per_step": ms_step,
                     "config": {"workload": "the reference's own decode (tiny-cuda-nn NetworkWithInputEncoding::inference, built unmodified from "
                                            f"/root/reference/tcnn for sm_100): the first-round batch of a {W}x{H} mode-5 frame of the {args.volume}^3 volume "
                                            f"in the reference marcher's layout ({xyz_frame.shape[0]} coordinates: 16 per live ray, step-major), the kernel "
                                            "sequence mode 5 issues per wavefront round.  The reference's marcher kernels cannot be built offline (OVR "
                                            "framework missing), so its frame rate is bounded above by this rate / samples per frame"},
                     "decode_samples_per_sec_frame_layout": v_frame, "decode_samples_per_sec_uniform_2p24": v_uniform,
                     "cpu_baseline": {"value": v, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "the reference's GPU decode on the same B200 (it has no CPU path); frame-layout and uniform batches, the faster one reported"},
                     "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(base)
        return
    # CPU port of the decode on the host cores
    m = O.ModelCfg()
    _, p16 = O.init_params(m, 1337)
    nb = 1 << 20
    xyz = np.random.default_rng(0).random((nb, 3), dtype=np.float32)
    O.decode(m, p16, xyz[:1 << 16])
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        O.decode(m, p16, xyz)
    dt = (time.perf_counter() - t0) / max(1, min(args.steps, 3))
    v = nb / dt
    base.update({"value": v, "ms_per_step": dt * 1e3, "config": {"workload": "CPU oracle port of the decode, 2^20 uniform samples per step"},
                 "cpu_baseline": {"value": v, "unit": UNIT, "cores": O.lib().orc_num_threads(), "kind": "port", "sample": "2^20 uniform samples per step"},
                 "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(base)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--volume", type=int, default=256)
    ap.add_argument("--frame", type=int, default=1024)
    ap.add_argument("--train-steps", type=int, default=600)
    ap.add_argument("--width", type=int, default=0, help="frame width (default --frame)")
    ap.add_argument("--height", type=int, default=0, help="frame height (default --frame)")
    ap.add_argument("--log2-hashmap", type=int, default=19, help="hash table size per level (config 5: 22)")
    ap.add_argument("--workload", default="render", choices=["render", "train", "interleaved"],
                    help="render = BASELINE configs[1] (the headline; --width 3840 --height 2160 --log2-hashmap 22 = configs[4]); "
                         "train = configs[2]/[3]: data-parallel training steps/s at --batch samples per rank")
    ap.add_argument("--batch", type=int, default=1 << 18, help="train workload: samples per rank per step")
    ap.add_argument("--out-of-core", action="store_true", help="train workload: the ground truth stays in a raw file, sampled through per-rank slab pools (configs[3])")
    ap.add_argument("--dp-mode", default="sharded", choices=["sharded", "allreduce"], help="train workload, N > 1: peer-memory optimizer or NCCL all-reduce")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--pipeline", type=int, default=0, help="render workload: frame slots inside the renderer for the device-resident timing (0 = 2 on one or two GPUs, 3 on four, 4 on eight)")
    ap.add_argument("--gather", default="comm", choices=["comm", "nccl"],
                    help="N > 1: comm = the library's communicator (peer stores from the compositing kernels); nccl = torch.distributed gather of padded strips (comparison)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    isolate_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    elif args.workload == "interleaved":
        if int(os.environ.get("RANK", "0")) == 0:
            run_interleaved(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
