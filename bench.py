#!/usr/bin/env python
"""bench.py -- GCUPS (incl. traceback) of the B200 alignment hot path, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c1] [--pairs P]
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

Workloads are BASELINE.json's configs.  The default is config 4
("batch of 1M short protein pairs, local alignment, sharded across 1/2/4/8 B200"):
it is the configuration the metric's 1/2/4/8-GPU quoting applies to that fits one
GPU; a "step" is one pass of the hot path over the whole batch.  Default scaling is
STRONG, as BASELINE words it: ONE batch of --pairs pairs, cut into cell-balanced
contiguous ranges (sa_partition_batch), rank r aligns range r, no data-path
collective.  --scaling weak gives every rank its own --pairs pairs instead.
c1/c2/c3 are the single-pair configs (replicas only at N>1).  After the main
measurement the line also carries config 5 (`c5`: one long global alignment as
column slices over the N GPUs; at N=1 the same 1 000 000 x 950 793 pair through the
checkpointed linear-space traceback, its 250 GB of directions never held at once)
unless --c5 off.

  value  = whole-job GCUPS with inputs resident in HBM (sa_align_batch_device /
           sa_align_device on torch's stream), fill + device traceback + string emission;
  e2e    = the same metric through the host-buffer C ABI (sa_align_batch / sa_align):
           pinned host inputs -> H2D -> kernels -> D2H of results and strings, per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sequence-alignment-gpu_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402

BLOSUM62 = None


def load_matrix(name):
    mats = json.load(open(os.path.join(ROOT, "tests", "golden", "matrices.json")))
    return np.asarray(mats[name], np.int32)


def peaks():
    p = dict(hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            p = dict(hbm_gbs=float(json.load(open(f))["hbm_gbs"]), source="MEASURED_PEAKS.json")
        except Exception:
            pass
    # integer / DPX issue rate measured by csrc/microbench/pipe_peaks.cu on this pool's B200
    # (profiles/r02_pipe_peaks.jsonl: VIADDMNMX and VIADDMNMX.S16x2 issue at the same rate, 18.2-18.3 T lane-ops/s)
    dpx = dict(s32=18.25e12, s16x2=18.25e12)
    src = "fallback (148 SMs x 62.8 lane-ops/clk x 1.965 GHz)"
    for name in ("r02_pipe_peaks.jsonl", "r01_pipe_peaks.jsonl"):
        f = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(f):
            continue
        for line in open(f):
            try:
                d = json.loads(line)
            except Exception:
                continue
            if d.get("op") == "VIADDMNMX":
                dpx["s32"] = d["lane_ops_per_s_T"] * 1e12
            if d.get("op") == "VIADDMNMX.S16x2":
                dpx["s16x2"] = d["lane_ops_per_s_T"] * 1e12
        src = f"csrc/microbench/pipe_peaks.cu on B200 (profiles/{name})"
        break
    p["dpx_lane_ops_per_s"] = dpx
    p["dpx_source"] = src
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    power_w_max=max(power), samples=len(sm))


# ------------------------------------------------------------------------------------------ workloads
def c4_name(args, world):
    per = "in total, sharded over the ranks" if args.scaling == "strong" else "per GPU"
    return f"c4: {args.pairs} protein pairs {per}, ~300 aa, mutate.py-style partners, SW, BLOSUM62, gap 5"


def make_workload(args, rank, world=1):
    if args.workload == "c4":
        if args.scaling == "strong":
            # ONE batch (seed 2024); every rank builds it and keeps its cell-balanced range (sa_partition_batch)
            T, toff, P, poff = synth.synthetic_batch(args.pairs, seed=2024)
            if world > 1:
                from __graft_entry__ import load_package
                first = load_package().partition_batch(toff, poff, world)
                a, b = int(first[rank]), int(first[rank + 1])
                T, P = T[toff[a]:toff[b]].copy(), P[poff[a]:poff[b]].copy()
                toff, poff = (toff[a:b + 1] - toff[a]).copy(), (poff[a:b + 1] - poff[a]).copy()
        else:
            T, toff, P, poff = synth.synthetic_batch(args.pairs, seed=2024 + rank)
        return dict(kind="batch", mode=1, alpha=23, matrix=load_matrix("protein/blosum62.txt"), gap=5,
                    text=T, toff=toff, pattern=P, poff=poff,
                    cells=int(((toff[1:] - toff[:-1] + 1) * (poff[1:] - poff[:-1] + 1)).sum()),
                    name=c4_name(args, world))
    seqs = np.load(os.path.join(ROOT, "tests", "golden", "sequences.npz"))
    if args.workload == "c1":
        a, b = seqs["dna/NC_018874.txt"], seqs["dna/mutated_NC_018874.txt"]
        w = dict(mode=0, alpha=4, matrix=load_matrix("dna/blast.txt"), gap=5,
                 name="c1: data/dna NC_018874 x mutated_NC_018874, NW, blast.txt, gap 5")
    elif args.workload == "c2":
        a, b = seqs["protein/P33450.fasta"], seqs["protein/mutated_P33450.fasta"]
        w = dict(mode=1, alpha=23, matrix=load_matrix("protein/blosum62.txt"), gap=5,
                 name="c2: data/protein P33450 x mutated_P33450, SW, BLOSUM62, gap 5")
    elif args.workload == "c3":
        a, b = synth.synthetic_pair(args.length, 12345, 54321)
        w = dict(mode=0, alpha=4, matrix=load_matrix("dna/blast.txt"), gap=5,
                 name=f"c3: synthetic DNA {args.length} x ~{int(args.length * 0.952)} (mutate.py-style, seeds 12345/54321), NW, blast.txt, gap 5")
    else:
        raise SystemExit("unknown workload")
    t, p = (a, b) if len(a) >= len(b) else (b, a)
    w.update(kind="single", text=np.ascontiguousarray(t), pattern=np.ascontiguousarray(p),
             cells=(len(t) + 1) * (len(p) + 1))
    return w


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation (oracle/_ref, else the oracle port) on the host cores."""
    if rank != 0:
        return
    from oracle.oracle_py import Oracle, Reference
    w = make_workload(args, 0) if args.workload != "c4" else None
    cores = os.cpu_count() or 1
    kind = "reference" if Reference.available("O3") else "port"
    if args.workload == "c4":
        sample_pairs = min(args.pairs, args.ref_pairs)
        T, toff, P, poff = synth.synthetic_batch(sample_pairs, seed=2024)
        mat = load_matrix("protein/blosum62.txt")
        cells = int(((toff[1:] - toff[:-1] + 1) * (poff[1:] - poff[:-1] + 1)).sum())
        if kind == "reference":
            ref = Reference("O3")
            step = lambda: ref.align_batch(1, 23, mat, 5, T, toff, P, poff, nthreads=cores)
        else:
            orc = Oracle()
            cores = 1
            step = lambda: [orc.align(1, 23, mat, 5, T[toff[i]:toff[i + 1]], P[poff[i]:poff[i + 1]]) for i in range(sample_pairs)]
        sample = f"{sample_pairs} pairs of the c4 batch per step (seed 2024), alignSequenceCPU incl. traceback, one pair per thread"
        name = c4_name(args, world)
    else:
        t, p = w["text"], w["pattern"]
        lim = args.ref_length
        t, p = t[:lim], p[:lim]
        cells = (len(t) + 1) * (len(p) + 1)
        cores = 1
        if kind == "reference":
            ref = Reference("O3")
            step = lambda: ref.align(w["mode"], w["alpha"], w["matrix"], w["gap"], t, p)
        else:
            orc = Oracle()
            step = lambda: orc.align(w["mode"], w["alpha"], w["matrix"], w["gap"], t, p)
        sample = f"leading {len(t)} x {len(p)} sub-problem of the pair, alignSequenceCPU incl. traceback, 1 thread"
        name = w["name"]
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    gcups = cells * args.steps / dt / 1e9
    line = dict(metric="GCUPS incl. traceback", value=gcups, unit="GCUPS", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=dt / args.steps * 1e3, higher_is_better=True,
                scaling=args.scaling if args.workload == "c4" else "weak",
                vs_baseline=None, dtype="int32", data="synthetic" if args.workload in ("c3", "c4") else "reference data/ files",
                impl="reference",
                config=dict(workload=name, l2="n/a: host cores only, no GPU in this arm", cells_per_step=cells, rank0_cells_per_step=cells,
                            parallelism=f"{cores} host threads, one pair per thread" if args.workload == "c4" else "1 host thread",
                            sample_per_step=sample,
                            note="the CPU arm times a bounded sample of the workload per step (throughput is size-independent "
                                 "for independent pairs)"),
                cpu_baseline=dict(value=gcups, unit="GCUPS", cores=cores, kind=kind, sample=sample),
                e2e=dict(value=gcups, unit="GCUPS", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def cpu_baseline(args, w):
    from oracle.oracle_py import Oracle, Reference
    cores = os.cpu_count() or 1
    kind = "reference" if Reference.available("O3") else "port"
    if w["kind"] == "batch":
        k = min(len(w["toff"]) - 1, args.ref_pairs)
        toff, poff = w["toff"][:k + 1], w["poff"][:k + 1]
        cells = int(((toff[1:] - toff[:-1] + 1) * (poff[1:] - poff[:-1] + 1)).sum())
        if kind == "reference":
            ref = Reference("O3")
            t0 = time.perf_counter()
            ref.align_batch(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], toff, w["pattern"], poff, nthreads=cores)
            dt = time.perf_counter() - t0
            t1 = time.perf_counter()
            k1 = max(1, k // cores)
            ref.align_batch(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], toff[:k1 + 1], w["pattern"], poff[:k1 + 1], nthreads=1)
            dt1 = time.perf_counter() - t1
            cells1 = int(((toff[1:k1 + 1] - toff[:k1] + 1) * (poff[1:k1 + 1] - poff[:k1] + 1)).sum())
            return dict(value=cells / dt / 1e9, unit="GCUPS", cores=cores, kind=kind,
                        sample=f"first {k} pairs of the batch, unmodified alignSequenceCPU (-O3) incl. traceback, one pair per thread",
                        one_core_gcups=cells1 / dt1 / 1e9)
        orc = Oracle()
        k = min(k, 4000)
        t0 = time.perf_counter()
        for i in range(k):
            orc.align(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"][toff[i]:toff[i + 1]], w["pattern"][poff[i]:poff[i + 1]])
        dt = time.perf_counter() - t0
        cells = int(((toff[1:k + 1] - toff[:k] + 1) * (poff[1:k + 1] - poff[:k] + 1)).sum())
        return dict(value=cells / dt / 1e9, unit="GCUPS", cores=1, kind="port", sample=f"first {k} pairs, oracle/sa_oracle.c")
    t, p = w["text"][:args.ref_length], w["pattern"][:args.ref_length]
    cells = (len(t) + 1) * (len(p) + 1)
    eng = Reference("O3") if kind == "reference" else Oracle()
    t0 = time.perf_counter()
    eng.align(w["mode"], w["alpha"], w["matrix"], w["gap"], t, p)
    dt = time.perf_counter() - t0
    return dict(value=cells / dt / 1e9, unit="GCUPS", cores=1, kind=kind,
                sample=f"leading {len(t)} x {len(p)} sub-problem, alignSequenceCPU incl. traceback, 1 thread")


def verify_batch(w, outs, sample, seed, nthreads):
    """Field-by-field and string-by-string check of `sample` random pairs of each result set in `outs` against the
    unmodified reference's alignSequenceCPU (oracle/_ref; the C restatement when it did not travel)."""
    from oracle.oracle_py import Oracle, Reference
    chk = Reference("O3") if Reference.available("O3") else Oracle()
    N = len(w["toff"]) - 1
    idx = np.unique(np.random.default_rng(seed).integers(0, N, min(sample, N))).astype(np.uint64)
    bad = 0
    for o in outs:
        b, _ = chk.check_batch(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], w["toff"], w["pattern"], w["poff"], o,
                               idx=idx, nthreads=nthreads)
        bad += b
    return len(idx) * len(outs), bad, "reference" if isinstance(chk, Reference) else "port"


def verify_single(args, w, got):
    """c1/c2: the whole Response against the reference CPU path; c3: the reference's known answer for the seeded pair
    (tests/golden) plus the re-score and spelling properties (the CPU path needs 9.5 GB and minutes)."""
    from oracle.oracle_py import Oracle, Reference
    if args.workload == "c3":
        orc = Oracle()
        ok = orc.rescore(got.aligned_text, got.aligned_pattern, w["alpha"], w["matrix"], w["gap"]) == got.score
        if args.length == 100_000:
            ok = ok and (got.score, got.aln_len, got.start_text, got.start_pattern) == (399463, 100254, 0, 0)
        return 1, int(not ok), "known answer of the reference CPU path for this seeded pair + re-score"
    eng = Reference("O3") if Reference.available("O3") else Oracle()
    want = eng.align(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], w["pattern"])
    return 1, int(got.key() != want.key()), "reference" if isinstance(eng, Reference) else "port"


def device_for_rank(local_rank, world):
    """CUDA device of a local rank.  The GPUs of an HGX box hang off two host bridges (devices 0..D/2-1 and D/2..D-1);
    the end-to-end path moves 1.24 GB per 1 M pairs through the host, so a run on fewer ranks than devices spreads its
    ranks over both halves (0, D/2, 1, D/2+1, ...) instead of filling the first one.  Measured on 8 x B200: four ranks on
    devices 0-3 copy at 11.3 + 12.0 GB/s each, two ranks at 17.6 + 18.7 GB/s.  NVLink is all-to-all (NV18), so the order
    does not matter to the linked slices of config 5.  SA_BENCH_DEVICE_ORDER=linear keeps device = local rank."""
    try:
        import torch
        ndev = torch.cuda.device_count()
    except Exception:
        return local_rank
    if os.environ.get("SA_BENCH_DEVICE_ORDER") == "linear" or ndev < 4 or ndev % 2 or world >= ndev or local_rank >= ndev:
        return local_rank
    half = ndev // 2
    return (local_rank // 2) + half * (local_rank % 2)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    sa = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    def measure():
        al = sa.Aligner(local_rank)
        w = make_workload(args, rank, world)
        tstream = torch.cuda.Stream(device=dev)           # the kernels are launched (and timed) on this stream
        torch.cuda.set_stream(tstream)
        stream = tstream.cuda_stream
        pk = peaks()

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        if w["kind"] == "batch":
            N = len(w["toff"]) - 1
            arena = int(w["toff"][-1] + w["poff"][-1])
            max_n = int((w["toff"][1:] - w["toff"][:-1]).max())
            max_m = int((w["poff"][1:] - w["poff"][:-1]).max())
            # pinned host copies (the e2e path reads these), device-resident copies (the `value` path)
            hT = torch.from_numpy(w["text"]).pin_memory(); hP = torch.from_numpy(w["pattern"]).pin_memory()
            hto = torch.from_numpy(w["toff"]).pin_memory(); hpo = torch.from_numpy(w["poff"]).pin_memory()
            dT, dP, dto, dpo = (x.to(dev, non_blocking=True) for x in (hT, hP, hto, hpo))
            dres = torch.zeros(N * 4, dtype=torch.int64, device=dev)
            daoff = torch.zeros(N, dtype=torch.int64, device=dev)
            doT = torch.empty(arena, dtype=torch.uint8, device=dev)
            doP = torch.empty(arena, dtype=torch.uint8, device=dev)
            hout = dict(results=torch.zeros(N * 4, dtype=torch.int64).pin_memory().numpy().view(sa.RESULT_DTYPE),
                        aln_off=torch.zeros(N, dtype=torch.int64).pin_memory().numpy().view(np.uint64),
                        aligned_text=torch.empty(arena, dtype=torch.uint8).pin_memory().numpy(),
                        aligned_pattern=torch.empty(arena, dtype=torch.uint8).pin_memory().numpy())

            def dev_step():
                al.align_batch_device(w["mode"], w["alpha"], w["matrix"], w["gap"], N, dT.data_ptr(), dto.data_ptr(),
                                      dP.data_ptr(), dpo.data_ptr(), dres.data_ptr(), daoff.data_ptr(), doT.data_ptr(),
                                      doP.data_ptr(), arena, max_n, max_m, stream=stream)

            def e2e_step():
                al.align_batch(w["mode"], w["alpha"], w["matrix"], w["gap"], hT.numpy(), hto.numpy(), hP.numpy(), hpo.numpy(), out=hout)

            h2d = int(hT.numel() + hP.numel() + 16 * (N + 1))
            d2h = int(32 * N + 8 * N + 2 * arena)
            l2_note = f"inputs {h2d / 1e6:.0f} MB and direction workspace >> 126 MB L2 (no flush needed)"
        else:
            n, m = len(w["text"]), len(w["pattern"])
            hT = torch.from_numpy(w["text"]).pin_memory(); hP = torch.from_numpy(w["pattern"]).pin_memory()
            dT, dP = hT.to(dev), hP.to(dev)
            doT = torch.empty(n + m, dtype=torch.uint8, device=dev)
            doP = torch.empty(n + m, dtype=torch.uint8, device=dev)
            dres = torch.zeros(4, dtype=torch.int64, device=dev)
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

            def dev_step():
                flush.fill_(1)          # evict L2 between timed iterations (inputs are smaller than L2)
                al.align_device(w["mode"], w["alpha"], w["matrix"], w["gap"], dT.data_ptr(), n, dP.data_ptr(), m,
                                doT.data_ptr(), doP.data_ptr(), dres.data_ptr(), stream=stream)

            def e2e_step():
                al.align(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], w["pattern"])

            h2d, d2h = n + m, 2 * (n + m) + 32
            l2_note = "L2 flushed between iterations (256 MiB fill)"

        # ---- device-resident timing (`value`) ----
        for _ in range(args.warmup):
            dev_step()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fill_us = tb_us = 0.0
        launches = 0
        ev0.record()
        per_step = []
        for _ in range(args.steps):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            dev_step()
            b.record()
            per_step.append((a, b))
        ev1.record()
        barrier()
        clocks = sampler.stop()
        t = al.timing()                     # kernel events of the LAST step
        launches = t["kernel_launches"]
        # kernel durations for the roofline: one extra step with the fill/traceback overlap switched off,
        # so that the CUDA events bracket the fill kernels alone (same kernels, same inputs)
        os.environ["SA_BATCH_PIPELINE"] = "0"
        dev_step()
        barrier()
        t = al.timing()
        os.environ.pop("SA_BATCH_PIPELINE", None)
        fill_us, tb_us = t["fill_us"], t["traceback_us"]
        if w["kind"] == "single":
            step_ms = [x.elapsed_time(y) for x, y in per_step]
            # the L2 flush is not part of the path: use the kernel events of the path itself
            dev_ms = (fill_us + tb_us) / 1e3
            total_ms = dev_ms * args.steps
        else:
            total_ms = ev0.elapsed_time(ev1)
            dev_ms = total_ms / args.steps
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        cc = torch.tensor([float(w["cells"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        total_ms_max = float(tt.item()); cells_all = float(cc.item())
        value = cells_all * args.steps / (total_ms_max * 1e-3) / 1e9

        # ---- end-to-end through the host-buffer C ABI ----
        for _ in range(max(1, args.warmup - 1)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        te = al.timing()
        et = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        # per-rank copy rates of the last e2e call (what bounds the multi-GPU end-to-end number: all ranks share one host)
        io = torch.tensor([float(te.get("h2d_bytes") or 0), float(te.get("d2h_bytes") or 0), e2e_s / args.steps], dtype=torch.float64, device=dev)
        ios = [torch.zeros_like(io) for _ in range(world)]
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_gather(ios, io)
        else:
            ios = [io]
        e2e_value = cells_all * args.steps / float(et.item()) / 1e9
        if te.get("d2h_bytes"):      # the batch entry point reports what its last call copied (strings are packed on the device)
            h2d, d2h = int(te["h2d_bytes"]), int(te["d2h_bytes"])

        # ---- verification of the outputs the two timed paths produced (after the timed regions) ----
        if w["kind"] == "batch":
            dev_out = dict(results=dres.cpu().numpy().view(sa.RESULT_DTYPE), aln_off=daoff.cpu().numpy().astype(np.uint64),
                           aligned_text=doT.cpu().numpy(), aligned_pattern=doP.cpu().numpy())
            vp, vbad, vkind = verify_batch(w, [dev_out, hout], max(1024, args.verify_pairs // world), 1 + rank,
                                           max(1, (os.cpu_count() or 1) // world))
            del dev_out
        else:
            vp, vbad, vkind = verify_single(args, w, al.align(w["mode"], w["alpha"], w["matrix"], w["gap"], w["text"], w["pattern"]))
        vv = torch.tensor([float(vp), float(vbad)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(vv, op=dist.ReduceOp.SUM)
        verified = dict(pairs=int(vv[0].item()), mismatches=int(vv[1].item()), against=vkind,
                        what=("device-resident and host-buffer outputs of the last timed steps, every Response field and both strings"
                              if w["kind"] == "batch" else "the Response of the host-buffer entry"))

        line = None
        if rank == 0:
            cells = float(w["cells"])
            fill_s = fill_us * 1e-6
            gcups_fill = cells / fill_s / 1e9 if fill_s > 0 else None
            # arithmetic mode of the kernel that ran: the batch kernels carry two pairs per register (s16x2) whenever
            # 4*max|H| fits 16 bits -- true for this workload -- the long-pair kernels are s32
            mode = "s16x2" if w["kind"] == "batch" else "int32"
            ops_per_cell = 1 if mode == "s16x2" else 2           # 2 DPX ops per cell in s32, 1 instruction per cell in s16x2
            dpx_peak = pk["dpx_lane_ops_per_s"]["s16x2" if mode == "s16x2" else "s32"]
            dpx_peak_gcups = dpx_peak / ops_per_cell / 1e9
            alg_bytes = 0.25 * cells + (int(hT.numel()) + int(hP.numel()))      # packed 2-bit directions written + sequences read
            hbm_ach = alg_bytes / fill_s / 1e9 if fill_s > 0 else None
            hbm_peak_gcups = pk["hbm_gbs"] / 0.25
            per_gpu_value = value / world
            traffic = None
            tf = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tf):
                try:
                    traffic = json.load(open(tf)).get(args.workload)
                except Exception:
                    traffic = None
            if traffic is not None and w["kind"] == "batch":
                traffic = traffic * (len(w["toff"]) - 1) / 1_000_000          # the capture is per 1 M pairs
            kname = ("batch_line16_kernel (all class launches of the step)" if w["kind"] == "batch"
                     else "tile_fill_kernel")
            line = dict(
                metric="GCUPS incl. traceback", value=value, unit="GCUPS", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=total_ms_max / args.steps, higher_is_better=True,
                scaling=(args.scaling if w["kind"] == "batch" else "weak"),
                vs_baseline=None, dtype=mode, data="synthetic" if args.workload in ("c3", "c4") else "reference data/ files",
                config=dict(workload=w["name"], l2=l2_note, cells_per_step=cells_all, rank0_cells_per_step=w["cells"],
                            devices=[device_for_rank(r, world) for r in range(world)],
                            parallelism=("one batch cut into cell-balanced contiguous ranges, one per rank, no collective" if w["kind"] == "batch" and args.scaling == "strong"
                                         else "every rank aligns its own batch, no collective" if w["kind"] == "batch" else "replicas only")),
                e2e=dict(value=e2e_value, unit="GCUPS", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=float(et.item()) / args.steps * 1e3,
                         per_rank=[dict(h2d_gbs=float(x[0]) / float(x[2]) / 1e9, d2h_gbs=float(x[1]) / float(x[2]) / 1e9,
                                        ms_per_step=float(x[2]) * 1e3) for x in ios]),
                gpu_launches=int(launches) * args.steps,
                clocks=clocks,
                verified=verified,
                # the BINDING roofline of this path is the DPX / integer-ALU issue rate, not HBM (report against the
                # slower bound); the HBM form the contract describes is kept inside as `hbm`
                roofline=dict(bound="dpx-alu", mode=mode, ops_per_cell=ops_per_cell,
                              achieved=(ops_per_cell * cells / fill_s) if fill_s > 0 else None, peak=dpx_peak, unit="lane-ops/s",
                              frac=(ops_per_cell * cells / fill_s / dpx_peak) if fill_s > 0 else None,
                              peak_source=pk["dpx_source"], kernel=kname,
                              kernel_ms_per_step=fill_us / 1e3, traceback_ms_per_step=tb_us / 1e3,
                              fill_only_gcups=gcups_fill, roofline_gcups=min(dpx_peak_gcups, hbm_peak_gcups),
                              frac_incl_traceback=per_gpu_value / min(dpx_peak_gcups, hbm_peak_gcups) if args.scaling == "weak" or world == 1 else None,
                              traffic=traffic,
                              traffic_source="dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the fill kernels "
                                             "(profiles/traffic.json), scaled to this step's pairs; not re-measured in this run",
                              hbm=dict(bound="hbm", achieved=hbm_ach, peak=pk["hbm_gbs"], unit="GB/s",
                                       frac=(hbm_ach / pk["hbm_gbs"]) if hbm_ach else None, peak_source=pk["source"],
                                       algorithmic_bytes_per_step=alg_bytes,
                                       note="0.25 B/cell of packed directions + the sequences; not the binding bound")),
            )
            if not args.no_cpu and world == 1:
                line["cpu_baseline"] = cpu_baseline(args, w)
        al.close()
        return line

    line = measure()

    # ---- config 5 beside it: one long global alignment as column slices over the same GPUs ----
    if args.c5 != "off":
        # free the batch buffers first: a slice of the 1 M x 0.95 M pair needs 125 GB of direction words at N = 2
        # (measure()'s device and pinned buffers went out of scope with it)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        from bench_c5 import run_c5
        length = int(args.c5) if args.c5 != "auto" else 1_000_000
        try:
            c5 = run_c5(sa, rank, world, local_rank, length, steps=2)          # best of two: the first call pays the allocations
        except SystemExit as e:          # does not fit: say so instead of failing the headline line
            c5 = dict(skipped=str(e)) if rank == 0 else None
        if rank == 0 and line is not None:
            line["c5"] = c5
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="c4: pairs in total (--scaling strong) or per GPU (weak)")
    ap.add_argument("--length", type=int, default=100_000, help="c3: text length")
    ap.add_argument("--ref-pairs", type=int, default=40_000, help="CPU sample size (pairs) for c4")
    ap.add_argument("--ref-length", type=int, default=30_000, help="CPU sample size (residues) for single-pair workloads")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="c4 over N GPUs: strong = ONE batch of --pairs pairs sharded over the ranks (BASELINE config 4), "
                         "weak = --pairs pairs on every rank")
    ap.add_argument("--verify-pairs", type=int, default=16384, help="pairs of each output set checked against the reference after the timed regions")
    ap.add_argument("--c5", default="auto", help="config 5 beside the main line: auto (1 000 000: column slices at N >= 2, the checkpointed traceback at N = 1), off, or a length")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = device_for_rank(int(os.environ.get("LOCAL_RANK", "0")), world)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
