#!/usr/bin/env python
"""bench.py -- 1080p macroblocks/s of the B200 reconstruction engine (BASELINE.json's metric).

A "step" is one pass of the hot path (reconstruct + in-loop filter + border, every picture) over one batch:
STREAMS independent instances of tests/golden/test_1920x1080.h264 per GPU, each with its own pre-parsed
work-list ("tape") and its own frame slots resident in HBM (BASELINE.json configs[2]: 512 looped streams on
one B200; configs[3] at N GPUs: 512 per GPU, statically sharded, no collective on the data path).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU)
  python bench.py --impl reference ...        the reference C decoder on the host cores (oracle/_ref)

Prints ONE JSON line on rank 0.  `value` = device-resident throughput; `e2e` = the same metric from host
bitstream bytes to host YUV frames through the C-ABI (host parse + H2D + GPU + D2H inside the timed region).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
STREAM = os.path.join(ROOT, "tests", "golden", "test_1920x1080.h264")
METRIC = "1080p macroblocks/s (H.264 Baseline macroblock reconstruction, bit-exact YUV)"
UNIT = "MB/s"
MB_REC_BYTES = 96   # D: work-list bytes per macroblock (intra macroblocks: + 2 in the order list)


def shard_streams(total, world, rank):
    """static shard: contiguous block of streams per rank; returns (first, count)"""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def logical_cpus():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def host_cores():
    """CPUs this process may actually use: the affinity mask, capped by the cgroup CPU quota (cpu.max = "quota period";
    a container with 128 visible CPUs and a quota of 16 is throttled, not sped up, by 128 busy threads)"""
    n = logical_cpus()
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return n


def _cpulist(txt):
    out = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b_ = part.partition("-")
        out.update(range(int(a), int(b_ or a) + 1))
    return out


def rank_cpu_set(local, world):
    """the CPUs this rank's parse / upload threads are pinned to: the usable CPUs that are local to its GPU (the NUMA node of
    the GPU's PCI device), shared evenly among the ranks whose GPUs sit on the same node, capped by the cgroup quota"""
    allowed = sorted(os.sched_getaffinity(0))
    quota = max(1, host_cores() // max(1, world))
    lists = []
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout
        for line in out.strip().splitlines()[:max(1, world)]:
            bus = line.strip().lower()
            bus = bus[-12:] if len(bus) > 12 else bus                      # 00000000:1B:00.0 -> 0000:1b:00.0
            lists.append(frozenset(_cpulist(open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read()) & set(allowed)))
    except Exception:
        lists = []
    if len(lists) <= local or not lists[local]:
        share = allowed[local::max(1, world)] or allowed
        return set(share[:quota]), f"{len(share[:quota])} of {len(allowed)} usable CPUs (GPU locality unknown)"
    peers = [r for r in range(len(lists)) if lists[r] == lists[local]]
    mine = sorted(lists[local])[peers.index(local)::len(peers)][:quota]
    return set(mine or allowed), f"{len(mine)} CPUs local to GPU {local} (its node has {len(lists[local])} usable, shared by {len(peers)} ranks)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for i, n in enumerate(names):
                if len(s) > 3 + i and s[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def gpu_busy(self):
        """mean of nvidia-smi's utilization.gpu over the samples (percent of time a kernel was running)"""
        u = [float(s[7]) for s in self.samples if len(s) > 7 and s[7].replace(".", "").isdigit()]
        return sum(u) / len(u) if u else None


def kernel_source_sha16():
    """identifies the build a profile belongs to: hash of the kernel sources"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "h264bsd_b200", "csrc", "engine")
    for n in sorted(os.listdir(d)):
        if n.endswith((".cuh", ".cu", ".hpp")):
            h.update(open(os.path.join(d, n), "rb").read())
    return h.hexdigest()[:16]


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_run(threads, seconds):
    """the reference C decoder's own decode loop on `threads` host threads (oracle/_ref/ref_loop)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_loop")
    if os.path.exists(exe):
        out = subprocess.run([exe, STREAM, str(threads), str(seconds)], capture_output=True, text=True, timeout=seconds * 20 + 120)
        r = json.loads(out.stdout.strip().splitlines()[-1])
        return {"value": r["mb_per_s"], "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{r['pics']} pictures ({r['mbs']} MB) of test_1920x1080.h264 decoded by the unmodified reference C "
                          f"(gcc -O3, posix flags) on {threads} host threads in {r['wall_s']:.1f} s"}, r["wall_s"], r["mbs"]
    # no compiled reference on this box: time the CPU oracle port (scalar, one thread)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    from h264bsd_b200.batch import ParsedStream
    ps = ParsedStream(open(STREAM, "rb").read())
    t0 = time.time()
    _oracle.oracle_run_tape(ps)
    dt = time.time() - t0
    mbs = ps.num_pics * ps.mbs_per_pic
    return {"value": mbs / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"one pass of test_1920x1080.h264 ({mbs} MB) through oracle/px_oracle.c, pixel path only, 1 thread"}, dt, mbs


def best_reference_threads(probe_seconds=1.5):
    """the thread count the reference runs fastest with here: the usable cores (quota-aware) or every logical CPU"""
    cands = sorted({host_cores(), logical_cpus()})
    if len(cands) == 1:
        return cands[0]
    best, best_v = cands[0], -1.0
    for c in cands:
        v = cpu_reference_run(c, probe_seconds)[0]["value"]
        if v > best_v:
            best, best_v = c, v
    return best


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = best_reference_threads()
    per_step = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_run(cores, 1.0)
    tot_mbs, tot_s = 0, 0.0
    cb = None
    for _ in range(args.steps):
        cb, wall, mbs = cpu_reference_run(cores, per_step)
        tot_mbs += mbs
        tot_s += wall
    value = tot_mbs / tot_s
    cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * tot_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"reference C decoder (oracle/_ref) looping test_1920x1080.h264 on {cores} host threads, "
                                   f"bounded sample of ~{per_step:.0f} s per step"},
            "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def algorithmic_bytes(ps):
    """per-picture algorithmic bytes of pass A (the fused MC + dequant/IDCT + add + write kernel), of the intra pass and of the
    in-loop filter: inter MB 768 + 32*nCoded + D, intra MB 384 + 32*nCoded + D (I_PCM: its 384 raw bytes are the
    12 'coded' blocks), D = the 96-byte record; a zero-motion copy 768 + 64 (pass A reads the two 32-byte sectors of the record
    that hold head, reference slots and first vector); deblock 768 + D per MB."""
    import numpy as np
    t = ps.ptr.contents
    area = np.ctypeslib.as_array(t.mbRecs, shape=(t.mbRecBytes,))
    pic_bytes = ps.mbs_per_pic * MB_REC_BYTES      # (a picture's records; a picture may be followed by filter-only records)
    recs = np.concatenate([area[p.mbRecOffset:p.mbRecOffset + pic_bytes] for p in ps.pics]).reshape(-1, MB_REC_BYTES)
    types = recs[:, 0]
    masks = recs[:, 4:8].copy().view("<u4")[:, 0] & 0x3FFFFFF
    pop = np.zeros(len(masks), np.int64)
    m = masks.astype(np.uint64)
    for b in range(26):
        pop += ((m >> np.uint64(b)) & np.uint64(1)).astype(np.int64)
    pop[types == 31] = 12
    inter = types <= 5
    pass_a = inter | (types == 31)
    mv0 = recs[:, 32:36].copy().view("<i2")
    copy = (types <= 1) & (masks == 0) & ((mv0[:, 0] | mv0[:, 1]) == 0)   # passAKernel: isCopy
    recon = np.where(inter, 768, 384) + 32 * pop + MB_REC_BYTES
    nmb = ps.mbs_per_pic
    per_pic_a = np.where(pass_a, np.where(copy, 768 + 64, recon), 0).reshape(-1, nmb).sum(axis=1)
    per_pic_b = np.where(pass_a, 0, recon + 2).reshape(-1, nmb).sum(axis=1)   # + 2: the macroblock's entry in the order list
    per_pic_deblock = np.full(ps.num_pics, (768 + MB_REC_BYTES) * nmb, np.int64)
    return per_pic_a, per_pic_b, per_pic_deblock, float(inter.mean()), float(pop.mean()), float(copy.mean()), float((pass_a & ~copy).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("B200_BENCH_STREAMS", "512")), help="streams per GPU")
    ap.add_argument("--e2e-streams", type=int, default=int(os.environ.get("B200_BENCH_E2E_STREAMS", "512")), help="streams per GPU in the end-to-end leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from h264bsd_b200.batch import Batch, ParsedStream

    data = open(STREAM, "rb").read()
    ps = ParsedStream(data)
    assert ps.status == 0 and ps.num_pics == 73
    total_streams = args.streams * world
    first, count = shard_streams(total_streams, world, rank)
    nmb = ps.mbs_per_pic
    mbs_per_step_rank = count * ps.num_pics * nmb
    per_pic_a_bytes, per_pic_intra_bytes, per_pic_deblock_bytes, inter_frac, coded_per_mb, copy_frac, other_frac = algorithmic_bytes(ps)

    b = Batch(count, ps.width_mbs, ps.height_mbs, ps.num_slots, device=local)
    b.upload(0, ps)
    b.replicate(0)  # every stream owns a distinct copy of the work-list in HBM
    b.sync()

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        b.run(0, ps.num_pics)
    b.sync()
    # parity gate inside the bench: last-pass output of first / last stream vs the golden md5, all streams equal
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "md5.json")))["test_1920x1080.h264"]
    last_slot = ps.pics[-1].curSlot
    ok = hashlib.md5(b.read_frame(0, last_slot).tobytes()).hexdigest() == gold["post_frame_md5"][-1]
    ok = ok and hashlib.md5(b.read_frame(count - 1, last_slot).tobytes()).hexdigest() == gold["post_frame_md5"][-1]
    ok = ok and b.compare_streams([last_slot] * count) == 0 and b.watchdog() == (0, 0) and b.idct_errors() == 0
    if not ok:
        print(json.dumps({"error": "parity check failed: output differs from the reference decoder", "rank": rank}), flush=True)
        sys.exit(1)

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = b.launches()
    work0 = b.deblock_work_mbs()
    barrier()
    b.sync()
    b.timer_start()
    for _ in range(args.steps):
        b.run(0, ps.num_pics)
    ms = b.timer_stop()
    barrier()
    launches = b.launches() - launches0
    deblock_work_frac = (b.deblock_work_mbs() - work0) / float(args.steps * mbs_per_step_rank)
    # per-kernel durations: the same K steps once more with CUDA events around every launch on the engine's stream.  In this
    # mode the engine issues the kernels of a picture one after the other (in the timed region above the copy pass and the
    # boundary strengths run on side streams next to pass A), so that every duration is that kernel's own.
    b.kernel_timing(True)
    b.timer_start()
    for _ in range(args.steps):
        b.run(0, ps.num_pics)
    ms_serial = b.timer_stop()
    sampler.stop_flag.set()
    stage_ms, stage_n = b.kernel_times()
    b.kernel_timing(False)
    if dist is not None:
        import torch
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([float(mbs_per_step_rank)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        mbs_per_step = float(tot.item())
        lt = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    else:
        mbs_per_step = float(mbs_per_step_rank)
    value = mbs_per_step * args.steps / (ms / 1000.0)

    # roofline of the dominant kernel (rank 0's device): algorithmic bytes per launch / mean launch duration
    peak, peak_src = measured_peak_gbs()
    gbs = lambda nbytes, msec: nbytes / max(1e-9, msec / 1000.0) / 1e9
    per_launch = lambda key: stage_ms[key] / max(1, stage_n[key])
    # pass A = the fused MC + dequant/IDCT + add + write path of SURVEY 8(d) for every inter / I_PCM macroblock of a picture: one
    # launch of passAKernel per picture (an IDR picture's launch finds nothing to do: every macroblock is intra)
    pass_a_bytes = float(per_pic_a_bytes.sum()) * count / max(1, ps.num_pics)
    # (the engine dispatches it as two instances -- one partition / several partitions -- timed one after the other)
    pass_a_ms = per_launch("recon") + per_launch("recon_multi")
    achieved = gbs(pass_a_bytes, pass_a_ms)
    # in-loop filter: pels (768 B) only of the macroblocks that have a non-zero boundary strength, record + strengths of all
    deb_bytes_per_launch = (768.0 * deblock_work_frac + MB_REC_BYTES + 17) * nmb * count
    deb_ms_per_launch = per_launch("deblock") + per_launch("strength")
    step_ms = max(1e-9, sum(stage_ms.values()))
    roof = {"bound": "hbm", "kernel": "passAKernel (fused MC + dequant/IDCT + add + write of every inter / I_PCM macroblock, zero-motion copies "
                                      "included; one launch per picture over all streams)", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": pass_a_bytes, "ms_per_launch": pass_a_ms,
            "share_of_step": (stage_ms["recon"] + stage_ms["recon_multi"]) / step_ms,
            "ms_per_launch_instances": {"one_partition_and_copies": per_launch("recon"), "several_partitions": per_launch("recon_multi")},
            "serialized_step_ms": ms_serial / args.steps,
            "macroblock_mix": {"zero_motion_copies": copy_frac, "other_inter_and_pcm": other_frac, "intra": 1.0 - copy_frac - other_frac},
            "other_kernels": {"strengthKernel + deblockKernel": {"achieved_gbs": gbs(deb_bytes_per_launch, deb_ms_per_launch), "ms_per_launch": deb_ms_per_launch,
                                                                 "frac": gbs(deb_bytes_per_launch, deb_ms_per_launch) / peak,
                                                                 "macroblocks_with_work": deblock_work_frac,
                                                                 "share_of_step": (stage_ms["deblock"] + stage_ms["strength"]) / step_ms},
                              "reconIntraKernel": {"ms_per_launch": per_launch("recon_intra"), "share_of_step": stage_ms["recon_intra"] / step_ms,
                                                   "achieved_gbs": gbs(float(per_pic_intra_bytes.mean()) * count, per_launch("recon_intra"))},
                              "borderKernel": {"ms_per_launch": per_launch("border"), "share_of_step": stage_ms["border"] / step_ms}}}
    # config 5 (BASELINE.json): the YUV -> ARGB output kernel, one 1080p frame of every stream per launch, 5.5 bytes per pel
    conv_reps = 3
    conv_ms = b.convert_bench_all(last_slot, 1, conv_reps) / conv_reps
    conv_bytes = count * (ps.width_mbs * 16) * (ps.height_mbs * 16) * 5.5
    roof["other_kernels"]["convertFrameKernel"] = {"achieved_gbs": gbs(conv_bytes, conv_ms), "ms_per_launch": conv_ms,
                                             "frac": gbs(conv_bytes, conv_ms) / peak,
                                             "note": "BGRA of the last output frame of all streams in one launch; not part of the timed step"}
    # DRAM bytes of one passAKernel launch from the committed `ncu --set full` capture, scaled to this run's stream count; only
    # a capture of THIS build counts (the file names the source hash of the kernels it was taken with)
    prof = os.path.join(ROOT, "profiles", "r02_passA_traffic.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            if pj.get("kernel_source_sha16") == kernel_source_sha16():
                roof["traffic"] = pj["dram_bytes_per_launch_per_stream"] * count
                roof["traffic_source"] = pj.get("source")
            else:
                roof["traffic_source"] = "no capture of this build (profiles/r02_passA_traffic.json is of another one)"
        except Exception:
            pass

    # ---- end to end through the C-ABI with host buffers: every rank decodes `ne` streams from bitstream bytes to host
    # frames at the same time (they share the host cores); value = all streams / slowest rank.
    # Two batches alternate: while the GPU replays pass i out of one of them and its pictures travel to the host, the host
    # threads parse the bitstreams of pass i+1 and upload their work-lists into the other (h264bsdB200BatchParseUploadBegin).
    e2e = None
    if not args.no_e2e:
        from h264bsd_b200 import _lib
        L = _lib.load()
        cpus, cpu_note = rank_cpu_set(local, world)
        os.sched_setaffinity(0, cpus)           # the parse / upload threads (and their page-locked tapes) stay on the GPU's node
        ne = max(1, min(args.e2e_streams if args.e2e_streams > 0 else 512, count))
        threads = max(1, min(ne, len(cpus)))
        b.close()
        batches = [Batch(ne, ps.width_mbs, ps.height_mbs, ps.num_slots, device=local) for _ in range(2)]
        pool = L.h264bsdB200ParseUploadPoolCreate(threads)
        fb = ps.frame_bytes
        host_out = [L.h264bsdB200HostAlloc(fb * ne) for _ in range(2)]   # page-locked landing zones, double buffered by picture
        assert pool and all(host_out)
        bits = (C.c_uint8 * len(data)).from_buffer_copy(data)             # the bitstream bytes every stream decodes
        bufs = [bits] * ne
        phase = {"parse_upload_wait": 0.0, "issue": 0.0, "gpu_and_d2h_wait": 0.0}

        def gpu_side(eb):
            for k in range(ps.num_pics):
                eb.decode_picture(k)                                 # GPU: reconstruct + in-loop filter + border
                eb.read_picture_all(k, host_out[k & 1], fb)          # D2H: picture k of every stream, de-stripped, page-locked

        for eb in batches:                                            # warm-up passes (allocations, page-locking)
            eb.parse_upload_wait(eb.parse_upload_begin(pool, bufs))
            gpu_side(eb)
            eb.sync()
        barrier()
        # The timed window is the pipeline in steady state: `reps` GPU passes (decode + D2H of every picture) and, overlapping
        # them, the `reps` parse + upload passes that feed the following ones; the very first parse, which nothing can overlap,
        # primes the pipeline before the clock starts (cold_start_value below includes it).
        reps = max(3, min(args.steps, 8))
        tc = time.time()
        tok = batches[0].parse_upload_begin(pool, bufs)
        batches[0].parse_upload_wait(tok)
        busy = ClockSampler(local)
        busy.start()
        h2d0, d2h0 = sum(x.h2d_bytes() for x in batches), sum(x.d2h_bytes() for x in batches)
        t0 = time.time()
        for i in range(reps):
            cur, nxt = batches[i & 1], batches[(i + 1) & 1]
            ti = time.time()
            gpu_side(cur)                                             # pass i: GPU + D2H, asynchronous
            tok = nxt.parse_upload_begin(pool, bufs)                  # pass i+1: parse + H2D on the host threads, overlapping it
            tg = time.time()
            cur.sync()
            te = time.time()
            nxt.parse_upload_wait(tok)
            tw = time.time()
            phase["issue"] += tg - ti
            phase["gpu_and_d2h_wait"] += te - tg
            phase["parse_upload_wait"] += tw - te
        t1 = time.time()
        busy.stop_flag.set()
        dt = (t1 - t0) / reps
        dt_cold = (t1 - tc) / reps             # the same passes with the priming parse counted in
        h2d = (sum(x.h2d_bytes() for x in batches) - h2d0) // reps
        d2h = (sum(x.d2h_bytes() for x in batches) - d2h0) // reps
        last = np.ctypeslib.as_array(C.cast(host_out[(ps.num_pics - 1) & 1], C.POINTER(C.c_uint8)), shape=(fb * ne,))
        ok2 = all(hashlib.md5(last[i * fb:(i + 1) * fb].tobytes()).hexdigest() == gold["post_frame_md5"][-1] for i in sorted({0, ne // 2, ne - 1}))
        ok2 = ok2 and all(x.watchdog() == (0, 0) and x.idct_errors() == 0 for x in batches)
        if dist is not None:
            import torch
            tt = torch.tensor([dt, dt_cold], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt, dt_cold = float(tt[0].item()), float(tt[1].item())
        e2e = {"value": world * ne * ps.num_pics * nmb / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "streams_per_gpu": ne, "host_threads_per_gpu": threads, "host_cpus": cpu_note,
               "bit_exact": bool(ok2), "passes": reps, "seconds_per_pass": dt,
               "cold_start_value": world * ne * ps.num_pics * nmb / dt_cold, "gpu_busy_pct": busy.gpu_busy(),
               "phase_seconds_per_pass": {k_: round(v_ / reps, 4) for k_, v_ in phase.items()},
               "note": "host bitstream bytes -> host I420 frames through the C-ABI: every host thread parses a stream into its page-locked tape and "
                       "uploads the work-list; the GPU replays a pass while the host parses the next one into a second batch; every output "
                       "picture is de-stripped on the GPU and copied into page-locked host memory.  The timed window holds `passes` parse + upload "
                       "passes and `passes` GPU + D2H passes of the running pipeline; cold_start_value also counts the first parse, which "
                       "nothing overlaps"}
        L.h264bsdB200ParseUploadPoolDestroy(pool)
        for hp in host_out:
            L.h264bsdB200HostFree(hp)
        for eb in batches:
            eb.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cores = best_reference_threads()
    cb, _, _ = cpu_reference_run(cores, args.cpu_seconds)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"{args.streams} looped test_1920x1080.h264 streams per GPU (BASELINE.json configs[2]; 73 pictures, "
                               f"{nmb} MB each), pre-parsed work-lists and frame slots resident in HBM, one private copy per stream",
                   "streams_per_gpu": args.streams, "pictures_per_step": ps.num_pics, "mb_per_step": mbs_per_step,
                   "mb_record_bytes": MB_REC_BYTES, "inter_mb_fraction": inter_frac, "coded_blocks_per_mb": coded_per_mb,
                   "l2": "inputs (work-lists + frames, tens of GB) far exceed the 126 MB L2; no explicit flush",
                   "sharding": "static contiguous stream blocks per rank, no data-path collective"},
        "roofline": roof, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
