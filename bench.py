#!/usr/bin/env python
"""bench.py -- mel frames/s of the FCL-taco2 inference hot path on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus 1 --steps 10 --warmup 3                       # our arm
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                                 # the reference's CPU path (oracle port)

A "step" = one pass of the whole hot path (encoder -> predictors -> length regulator -> decoder
-> postnet) over one synthetic LJSpeech-shaped batch. `value` = frames of all ranks / device time
(inputs resident in HBM); `e2e` = the same through model.inference_stream() with host buffers
(host planning, H2D of the inputs and D2H of the mels inside the timed region; the D2H of a batch overlaps
the compute of the next one, as in the decode driver).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout. Libraries chat on fd 1 (NCCL prints its version banner there), so fd 1 is
# pointed at stderr for the whole run and the JSON line is written to the saved, real stdout at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

from fcl_taco2_b200 import synth, hparams            # noqa: E402

# algorithmic work per unit (SURVEY.md 8(d) / BASELINE.md 3), MAC counts of the reference formulation
MACS = {
    "S": dict(decoder_row_step=1_438_720, postnet_frame=348_160, front_phoneme=3_593_856),
    "T": dict(decoder_row_step=15_941_632, postnet_frame=4_341_760, front_phoneme=8_611_968),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    In-process NVML (nvidia_ml_py) on a thread: a few light queries every 10 ms. An `nvidia-smi -lms` subprocess
    was measured to stall kernel launches for tens of ms when a poll (or its start-up) lands inside a timed step;
    it is only the fallback when NVML cannot be imported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.samples = index, None, [], []
        self.nvml, self.handle, self.thread, self.run_flag, self.first = None, None, None, False, 0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.run_flag = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _loop(self):
        n = self.nvml
        while self.run_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, reasons))
            except Exception:
                pass
            time.sleep(0.01)

    def ready(self):
        return bool(self.samples) or bool(self.lines) or (self.nvml is None and self.proc is None)

    def mark(self):
        """Start of the timed region (the sampler is started before the warm-up)."""
        self.first = len(self.samples) if self.nvml is not None else len(self.lines)

    def stop(self):
        if self.nvml is not None:
            self.run_flag = False
            self.thread.join(timeout=1)
            n = self.nvml
            sel = self.samples[max(self.first - 1, 0):]
            names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in sel))
            mhz = [m for m, _ in sel if m > 0]
            return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sel), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.lines[max(self.first - 1, 0):]:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def workload(args, rank):
    """LJSpeech-shaped synthetic batch for this rank; different seed per rank (weak scaling)."""
    xs, ds = synth.synth_batch(args.batch, seed=args.seed + 1000 * rank, stress=args.stress,
                               fixed_len=500 if args.stress else None)
    return xs, ds


def cpu_port_time(kind, xs, ds, budget_s, seed):
    """Time the oracle port (the reference's algorithm on torch CPU, per-utterance loop like
    tts.py:655-674) on a bounded sample. -> (frames/s, n_utts, frames, seconds, threads)"""
    from oracle import restate          # CPU baseline leg: the only place bench.py touches oracle/
    torch.set_num_threads(os.cpu_count())
    sd = synth.random_state_dict(hparams.preset(kind), seed)
    drop = restate.Dropout(0.5, 1, native=True)       # torch's F.dropout, as the reference runs it
    with torch.no_grad():
        restate.inference(sd, torch.from_numpy(xs[0]), dur=ds[0], dropout=drop, fast_lstm=True)     # warm-up
        t0 = time.perf_counter()
        frames = n = 0
        for x, d in zip(xs, ds):
            out = restate.inference(sd, torch.from_numpy(x), dur=d, dropout=drop, utt_index=n, fast_lstm=True)
            frames += out.shape[0]
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return frames / dt, n, frames, dt, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    xs, ds = workload(args, 0)
    vals, last = [], None
    for it in range(args.warmup + args.steps):
        last = cpu_port_time(args.model, xs, ds, budget_s=max(2.0, 60.0 / max(1, args.steps + args.warmup)), seed=args.seed)
        if it >= args.warmup:
            vals.append(last)
    fps = sum(v[2] for v in vals) / sum(v[3] for v in vals)
    sample = f"first {vals[-1][1]} utterances ({vals[-1][2]} frames) of the batch-{args.batch} workload per step, per-utterance loop"
    line = {
        "metric": "mel frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(v[3] for v in vals) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": vals[-1][4], "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is Python and cannot travel to the GPU box: this is oracle/restate.py, the op-for-op "
                "torch-CPU restatement validated against the reference in the build container",
    }
    emit(line)


def config_dict(args, world):
    return {"workload": f"FCL-taco2-{args.model} batched inference, batch {args.batch} synthetic LJSpeech-shaped "
                        f"utterances per GPU" + (" (500-phoneme stress)" if args.stress else ""),
            "model_size": args.model, "per_gpu_batch": args.batch, "global_batch": args.batch * world,
            "precision": args.precision, "prenet_dropout": args.dropout, "forced_durations": True,
            "parallelism": f"utterance-sharded x{world}, no hot-path collective, final mel gather to rank 0",
            "l2": "flushed between timed steps (256 MiB write, outside the per-step events)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="S", choices=["S", "T"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--stress", action="store_true")
    ap.add_argument("--latency-utts", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--pair", type=int, default=-1, help="cta_group::2 decoder: 1 on, 0 off, -1 engine default")
    ap.add_argument("--e2e-planner", type=int, default=0, help="inference_stream: plan batches in a helper thread (1) or inline (0)")
    ap.add_argument("--dropout", type=float, default=0.5, help="prenet dropout rate (reference default 0.5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from fcl_taco2_b200 import model as M, plan as planmod, dist as fdist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fcl_nccl_%h_%p.log")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    m = M.from_preset(args.model, seed=args.seed, device=dev, precision=args.precision)
    m.set_prenet_dropout(rate=args.dropout, seed=1)
    eng = m.engine()
    if args.pair >= 0:
        eng.use_pair = bool(args.pair)
    xs, ds = workload(args, rank)
    pl = planmod.make_plan(xs, ds)
    n_frames = int(sum(int(d.sum()) for d in ds))
    n_rows = pl.n_rows
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dinp, h2d_bytes = eng.upload(pl)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # groups of utterances whose mels are sent while the next group's postnet runs; splitting the fused postnet costs
    # ~0.2 ms, which only pays once the gather (bound by rank 0's ingress, (N-1) x 185 MB) is longer than that
    K_CHUNKS = 1 if world <= 2 else 4
    gather = None
    if world > 1:   # forced durations: every rank's output chunk boundaries are known (and exchanged) before the pass
        from fcl_taco2_b200.plan import output_chunks
        per_utt = np.add.reduceat(pl.dur.astype(np.int64), pl.utt_off[:-1].astype(np.int64))
        ch = output_chunks(np.concatenate([[0], np.cumsum(per_utt)]), K_CHUNKS)
        bounds = [c[2] for c in ch] + [ch[-1][3]]
        bounds += [bounds[-1]] * (K_CHUNKS + 1 - len(bounds))
        gather = fdist.ChunkedGather(bounds, m.odim, dev)

    def one_step(timed):
        flush.fill_(rank + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.stage_events = [] if timed else None
        e0.record()
        if world > 1:
            res = eng.run_uploaded(pl, dinp, m.hp.zoneout_rate, args.dropout, 1, out_chunks=K_CHUNKS, chunk_cb=gather.on_chunk)
            with eng.stage("gather_tail"):
                gather.finish()            # transfers were started chunk by chunk during the postnet
        else:
            res = eng.run_uploaded(pl, dinp, m.hp.zoneout_rate, args.dropout, 1)
        e1.record()
        return e0, e1, None, eng.stage_events, None    # results are dropped: holding K outputs alive would force a
                                                       # fresh cudaMalloc of the output buffer inside every timed step

    import gc
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step(False)
    if rank == 0:
        t_wait = time.time()
        while not sampler.ready() and time.time() - t_wait < 3.0:
            time.sleep(0.02)                       # the sampler is up and polling before the timed region starts
    gc.collect()
    gc.disable()            # a host GC pause between launches would show up as GPU idle time inside the events
    barrier()
    sampler.mark()
    launches0 = eng.launches
    evs = [one_step(True) for _ in range(args.steps)]
    barrier()
    launches = eng.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [e0.elapsed_time(e1) for e0, e1, *_ in evs]
    total_ms = float(sum(step_ms))
    stage_ms = {}
    for _, _, _, se, _ in evs:
        for name, a, b in se:
            stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b)
    eng.stage_events = None

    # ---- e2e through the public API: host ids/durations in, host mels out. model.inference_stream() is the call a
    # decode driver makes: every batch is planned on the host, uploaded, decoded, and its mels are copied to pinned host
    # memory on a copy stream while the next batch computes. The timed region covers K whole batches, from the first
    # upload to the arrival of the last batch's mels on the host (L2 flushed before every batch, inside the region).
    def batches(k):
        for _ in range(k):
            yield {"xs": xs, "durs": ds}
    flush_l2 = lambda: flush.fill_(1)
    for _ in m.inference_stream(batches(6), before_batch=flush_l2, planner_thread=bool(args.e2e_planner)):
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    got = 0
    for outs in m.inference_stream(batches(args.steps), before_batch=flush_l2, planner_thread=bool(args.e2e_planner)):      # the generator returns a batch only when it is on the host
        got += len(outs)
    e1.record()
    barrier()
    assert got == args.steps * len(xs)
    e2e_ms = float(e0.elapsed_time(e1))

    # ---- reduce over ranks: max time, sum frames
    t = torch.tensor([total_ms, e2e_ms, float(n_frames), float(n_rows), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms = float(tmax[0]), float(tmax[1])
        all_frames, all_rows, all_launches = float(tsum[2]), float(tsum[3]), int(tsum[4])
    else:
        all_frames, all_rows, all_launches = float(n_frames), float(n_rows), launches

    if rank == 0:
        pk = peaks()
        macs = MACS[args.model]
        dec_ms = stage_ms.get("decoder_loop", 0.0) / args.steps
        dec_flops = 2.0 * macs["decoder_row_step"] * n_frames
        achieved = dec_flops / (dec_ms * 1e-3) / 1e12 if dec_ms > 0 else 0.0
        value = all_frames * args.steps / (total_ms * 1e-3)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "decoder_traffic.json")))
            traffic = tj.get(f"{args.model}_{args.batch}_{args.precision}", {}).get("bytes")
        except Exception:
            pass
        line = {
            "metric": "mel frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": config_dict(args, world),
            "frames_per_step": all_frames, "phoneme_rows_per_step": all_rows,
            "clocks": clocks,
            "e2e": {"value": all_frames * args.steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(n_frames * m.odim * 4),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": all_launches,
            "roofline": {"kernel": "decoder step loop (fcl_decoder_*), rank 0", "bound": "tensor", "achieved": achieved,
                         "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"], "traffic": traffic,
                         "peak_source": pk["src"] + " bf16_tflops_sustained", "ms_per_launch": dec_ms,
                         "algorithmic_flops_per_launch": dec_flops,
                         "note": "algorithmic FLOPs = 2 x MAC per useful row-step (reference formulation, nothing hoisted) x frames"},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        }
        # the HBM-side view SURVEY 8(d) asks for next to the tensor roofline: algorithmic bytes / stage time, rank 0
        sec = []
        for name, key, nbytes in (("postnet (first input + last output only: 640 B/frame)", "postnet", 640.0 * n_frames),
                                  ("length regulator scan + frame map (8 B/phoneme + 20 B/frame)", None,
                                   8.0 * n_rows + 20.0 * n_frames)):
            ms = (stage_ms.get(key, 0.0) if key else stage_ms.get("len_reg", 0.0) + stage_ms.get("frame_map", 0.0)) / args.steps
            if ms > 0:
                gbs = nbytes / (ms * 1e-3) / 1e9
                sec.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                            "frac": gbs / pk["hbm"], "ms": ms})
        line["roofline_secondary"] = sec
        if not args.no_extra:
            line["p50_utt_latency_ms"] = latency_p50(m, args, dev)
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (torchrun pins OMP threads to 1)
            fps, n, fr, dt, thr = cpu_port_time(args.model, xs, ds, budget_s=15.0, seed=args.seed)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": thr, "kind": "port",
                                    "sample": f"first {n} utterances ({fr} frames) of the same workload, per-utterance "
                                              f"loop of oracle/restate.py (torch CPU fp32), {dt:.1f} s"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def latency_p50(m, args, dev):
    """p50 single-utterance latency (batch = 1), device-timed through the public API with host input."""
    xs, ds = synth.synth_batch(args.latency_utts, seed=args.seed + 77)
    lat = []
    for i, (x, d) in enumerate(zip(xs, ds)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.inference(torch.from_numpy(x), None, dur=d)
        e1.record()
        e1.synchronize()
        if i >= 5:
            lat.append(e0.elapsed_time(e1))
    return float(np.median(lat))


if __name__ == "__main__":
    main()
