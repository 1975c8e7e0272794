"""bench.py -- headline benchmark of the bzip2 block pipeline (BASELINE.json).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA, libb2bz.so)
  python bench.py --impl reference --gpus 1 ...            the reference's CPU path (oracle port)
  torchrun ... bench.py --gpus N ...                       one rank per GPU, weak scaling

A step = one bzip2 -9 encode of the workload (BASELINE configs[1]: 1 GiB synthetic ASCII per GPU,
numpy PCG64 seed 20260923, 94 printable bytes + newline).  `value` = whole-job MB/s (10^6 raw bytes
per second) with the input resident in HBM; `e2e` = the same through the host-buffer C ABI call
(b2_bzip2_compress: H2D + all kernels + D2H inside the timed region).  The roofline entry is for the
dominant kernel (k_radix_pass, the onesweep pass of the BWT suffix sort): algorithmic bytes per launch
over its CUDA-event time, against the measured HBM copy bandwidth.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 20260923
LEVEL = 9
METRIC = "bzip2_-9_encode_MBps"


def gen_ascii(nbytes, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    out = np.empty(nbytes, dtype=np.uint8)
    step = 1 << 26
    for o in range(0, nbytes, step):
        k = min(step, nbytes - o)
        a = g.integers(32, 127, size=k, dtype=np.uint8)
        a[a == 126] = 10
        out[o:o + k] = a
    return out


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (perf_counter) -- the sampler is started before the warm-up because
        nvidia-smi needs about a second to deliver its first line."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        inside = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        window = "timed region"
        if not inside:
            inside, window = [r for (_, r) in self.rows], "whole run (no sample fell into the timed region)"
        self.rows = inside
        self.window = window
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows), "window": self.window}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles():
    p = os.path.join(ROOT, "profiles", "radix_pass_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def cpu_sample_blocks(data, nblocks):
    bs = LEVEL * 100000 - 19
    return data[: min(len(data), nblocks * bs)]


def run_reference(args):
    """The reference's own CPU implementation of the path (its JavaScript cannot run here: no node in the
    image; this is the C restatement in oracle/, all host threads), on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 32))  # one 900k block (~60 MB of working set) per thread; more threads only thrash the host caches
    per_step_blocks = max(2, cores)  # one 900k block per core and step keeps the run in minutes
    mb = int(os.environ.get("B2_BENCH_MB", "1024"))
    data = gen_ascii(min(mb << 20, per_step_blocks * 900000 + 1000), SEED)
    sample = np.ascontiguousarray(cpu_sample_blocks(data, per_step_blocks))
    for _ in range(args.warmup):
        O.bzip2_compress(sample[: 2 * 900000], LEVEL, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.bzip2_compress(sample, LEVEL, threads=cores)
    dt = time.perf_counter() - t0
    val = sample.size * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "1 GiB synthetic ASCII (PCG64 seed %d), bzip2 -9 (900k blocks) encode" % SEED, "level": LEVEL,
                   "sample": "first %d blocks (%d bytes) per step" % (per_step_blocks, sample.size)},
        "cpu_baseline": {"value": val, "unit": "MB/s", "cores": cores, "kind": "port",
                         "sample": "%d x 900k blocks per step, %d threads (oracle/bz2_oracle.c, one block per thread)" % (per_step_blocks, cores)},
        "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference JS cannot execute in this image (no node); C restatement of its algorithm timed instead",
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mb", type=int, default=int(os.environ.get("B2_BENCH_MB", "1024")), help="MiB of input per GPU and step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from compressjs_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _native.lib()
    rc = L.b2_init(local)
    if rc:
        raise SystemExit("b2_init: " + _native.last_error())

    shard = args.mb << 20
    nbytes = shard * world
    # weak scaling: the job is ONE stream of world x shard bytes; every rank holds the input (the block
    # cutting scan needs it), encodes a contiguous range of blocks, and the fragments are gathered over NCCL.
    host = np.concatenate([gen_ascii(shard, SEED + r) for r in range(world)]) if world > 1 else gen_ascii(shard, SEED)
    pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = host
    d_in = pinned.cuda(non_blocking=False)
    cap = L.b2_bzip2_bound(nbytes)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda") if world == 1 else None
    out_n = C.c_size_t()
    from compressjs_b200 import sharded as SH
    state = {"comp": 0}

    def step_resident():
        if world == 1:
            rc = L.b2_bzip2_compress_dev(d_in.data_ptr(), nbytes, LEVEL, d_out.data_ptr(), cap, C.byref(out_n))
            if rc:
                raise SystemExit("compress_dev failed: " + _native.last_error())
            state["comp"] = out_n.value
            return _native.stats()
        out = SH.compress_file_sharded(d_in, LEVEL)
        st = _native.stats()
        if out is not None:
            state["comp"] = out.numel()
        return st

    def step_e2e(check=False):
        if world == 1:
            out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
            rc = L.b2_bzip2_compress(pinned.data_ptr(), nbytes, LEVEL, C.byref(out), C.byref(n))
            if rc:
                raise SystemExit("compress failed: " + _native.last_error())
            st = _native.stats()
            if check:  # untimed warm-up call: the host-buffer path must produce the resident path's stream
                got = torch.from_numpy(np.ctypeslib.as_array(out, (n.value,))).cuda()
                if n.value != state["comp"] or not torch.equal(got, d_out[:n.value]):
                    raise SystemExit("e2e stream differs from the HBM-resident stream")
            L.b2_free(out)
            return st, n.value
        d = pinned.cuda(non_blocking=True)          # H2D of the step's input
        torch.cuda.current_stream().synchronize()   # the library works on its own stream
        out = SH.compress_file_sharded(d, LEVEL)
        nn = 0
        if out is not None:
            nn = out.numel()
            if state.get("pinned_out") is None or state["pinned_out"].numel() < nn:
                state["pinned_out"] = torch.empty(nn + (nn >> 3), dtype=torch.uint8, pin_memory=True)
            state["pinned_out"][:nn].copy_(out, non_blocking=True)   # D2H of the step's result into pinned memory
            torch.cuda.current_stream().synchronize()
        return _native.stats(), nn

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident (HBM) arm ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    t0 = time.perf_counter()
    agg = {}
    dev_ms = 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        st = step_resident()
        dev_ms += st["ms_total"]
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    if world > 1:
        dev_ms = ev0.elapsed_time(ev1)  # includes the NCCL gather and the assembly on rank 0
    clocks = sampler.stop(t0, t0 + wall) if rank == 0 else None
    comp_bytes = state["comp"]

    # device time: max over ranks (events on the library's launching stream)
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = t.tolist()

    # ---- e2e arm: host buffers through the C ABI ----
    e2e_steps = max(1, min(args.steps, 3))
    step_e2e(check=True)
    barrier()
    t1 = time.perf_counter()
    e2e_comp = 0
    for _ in range(e2e_steps):
        _, e2e_comp = step_e2e()
    barrier()
    e2e_wall = time.perf_counter() - t1
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_wall = t.item()

    if rank == 0:
        total_raw = nbytes
        value = total_raw * args.steps / (dev_ms_max / 1e3) / 1e6
        peak, peak_src = peaks()
        radix_gbs = (agg["radix_bytes"] / 1e9) / (agg["ms_radix"] / 1e3) if agg.get("ms_radix") else 0.0
        bwt_gbs = (agg["bwt_bytes"] / 1e9) / (agg["ms_bwt"] / 1e3) if agg.get("ms_bwt") else 0.0
        tr = traffic_from_profiles()
        line = {
            "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "%d MiB synthetic ASCII per GPU (PCG64 seed %d+rank), bzip2 -9 (900k blocks) encode" % (args.mb, SEED),
                       "level": LEVEL, "bytes_per_gpu": shard, "total_bytes": nbytes, "blocks_per_gpu": int(agg["blocks"] // max(args.steps, 1)),
                       "l2": "inputs (%d MiB) larger than L2 (126 MB); no flush needed" % args.mb, "bwt_batch_blocks": int(os.environ.get("B2_BWT_BATCH", "296")),
                       "compressed_bytes": comp_bytes, "wall_ms_per_step": wall_ms_max / args.steps},
            "e2e": {"value": total_raw * e2e_steps / e2e_wall / 1e6, "unit": "MB/s", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": e2e_comp,
                    "steps": e2e_steps, "api": "b2_bzip2_compress (host pinned in, library-pinned out; upload in 64 MiB chunks and download per batch overlapped with the encode)" if world == 1 else
                    "sharded.compress_file_sharded (pinned host in on every rank, stream gathered to rank 0 over NCCL, D2H on rank 0)"},
            "gpu_launches": int(agg["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_radix_pass (BWT onesweep pass)", "achieved": radix_gbs, "peak": peak, "unit": "GB/s",
                         "frac": radix_gbs / peak if peak else None,
                         "traffic": ((tr or {}).get("dram_bytes_per_record") or 0) * (agg["radix_bytes"] / max(agg["radix_launches"], 1) / 16.0) or None,
                         "traffic_source": (tr or {}).get("source"),
                         "peak_source": peak_src, "launches": int(agg["radix_launches"]),
                         "algorithmic_bytes_per_launch": agg["radix_bytes"] / max(agg["radix_launches"], 1),
                         "avg_launch_ms": agg["ms_radix"] / max(agg["radix_launches"], 1),
                         "bwt_stage": {"achieved": bwt_gbs, "frac": bwt_gbs / peak if peak else None, "rounds": int(agg["bwt_rounds"] // max(args.steps, 1)),
                                       "ms_per_step": agg["ms_bwt"] / args.steps}},
            "stages_ms_per_step": {k: agg[k] / args.steps for k in ("ms_rle1", "ms_bwt", "ms_mtf", "ms_huff", "ms_pack", "ms_radix")},
            "clocks": clocks,
        }
        if world > 1:
            line["sharded_phases_ms_last_step_rank0"] = {k: round(v, 2) for k, v in SH.PHASES.items()}
        if world == 1:
            # decode leg: the stream just produced, HBM resident (b2_bzip2_decompress_dev) -- the second half of the metric
            d_dec = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            dn = C.c_size_t()
            comp = comp_bytes
            L.b2_bzip2_decompress_dev(d_out.data_ptr(), comp, 0, d_dec.data_ptr(), nbytes, C.byref(dn))
            dms, dst = 0.0, None
            dsteps = max(1, min(args.steps, 3))
            for _ in range(dsteps):
                rc = L.b2_bzip2_decompress_dev(d_out.data_ptr(), comp, 0, d_dec.data_ptr(), nbytes, C.byref(dn))
                if rc:
                    raise SystemExit("decompress_dev failed: " + _native.last_error())
                dst = _native.stats()
                dms += dst["ms_total"]
            ok = bool(torch.equal(d_dec[: dn.value], d_in)) and dn.value == nbytes
            line["decode"] = {"metric": "bzip2_-9_decode_MBps", "value": nbytes * dsteps / (dms / 1e3) / 1e6, "unit": "MB/s", "steps": dsteps,
                              "roundtrip_ok": ok, "ms_per_step": dms / dsteps,
                              "stages_ms": {k: dst[k] for k in ("ms_scan", "ms_hdec", "ms_unmtf", "ms_ibwt", "ms_unrle")}}
        if not args.no_cpu and world == 1:
            from oracle import oracle as O
            O.build()
            sample = np.ascontiguousarray(cpu_sample_blocks(host, 8))
            tc = time.perf_counter()
            zc = O.bzip2_compress(sample, LEVEL)
            dtc = time.perf_counter() - tc
            line["cpu_baseline"] = {"value": sample.size / dtc / 1e6, "unit": "MB/s", "cores": 1, "kind": "port",
                                    "sample": "first 8 x 900k blocks (%d bytes) of the same workload, oracle/bz2_oracle.c, 1 thread" % sample.size,
                                    "compressed_bytes": len(zc)}
            # yardsticks SURVEY.md section 8(d) asks for next to the port: libbz2 1.0.8 on one core (a different, much
            # cheaper table search: NOT bit-compatible) and the reference's own published single-thread figure
            import bz2
            tl = time.perf_counter()
            zl = bz2.compress(sample.tobytes(), LEVEL)
            line["cpu_baseline"]["libbz2_1thread_MBps"] = sample.size / (time.perf_counter() - tl) / 1e6
            line["cpu_baseline"]["libbz2_compressed_bytes"] = len(zl)
            line["cpu_baseline"]["reference_js_published_MBps"] = 0.0936   # README.md:70 of the reference (enwik8, node 0.8, 2013 laptop)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
