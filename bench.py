#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Blokus simulation core.

Workload (BASELINE.json configs[1]): batched uniform-random 4-player self-play, 65,536 concurrent
games from the empty board to terminal score per GPU.  One "step" = one such batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # CUDA arm (this repo)
  python bench.py --impl reference [...]                          # CPU arm: the reference's own Python engine
                                                                  # (oracle/_ref) on all host cores

For N > 1 the driver launches one rank per GPU with torch.distributed.run; games are sharded by global
game index (rank r plays games [r*G, (r+1)*G) of every step: weak scaling, no data-path collective).
Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GAMES_PER_GPU = 65536
SEED = 42                      # configs/stage3_selfplay_gpu.yaml:3 of the reference
METRIC = "random_playouts_per_sec"
UNIT = "playouts/s"
WORKLOAD = "uniform-random 4-player self-play, 65536 concurrent games/GPU, empty board to terminal score"

# Algorithmic constants of one playout (DESIGN.md "Roofline"): bytes that must cross HBM and the
# warp-level instructions the row-sliced algorithm issues (ncu smsp__inst_executed.sum / games,
# profiles/r01_playout_ncu.md).
ALG_BYTES_PER_PLAYOUT_DEVICE = 16            # int32 game_scores[4] written; no input (empty board start)
ALG_WARP_INSTR_PER_PLAYOUT = None            # filled from profiles/ when known; see _int_roofline


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # under load = samples in the upper half of what was seen (idle samples at the edges drop out)
        load = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference engine (oracle/blokus_oracle.c), all host threads
# ------------------------------------------------------------------------------------------------
def _cpu_playouts(n_games: int, threads: int, game_offset: int) -> float:
    """Plays n_games oracle games on `threads` threads; returns elapsed seconds."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import orc

    orc.lib()
    t0 = time.perf_counter()
    if threads <= 1:
        for g in range(n_games):
            orc.playout(SEED, game_offset + g)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:   # ctypes releases the GIL inside the C call
            list(ex.map(lambda g: orc.playout(SEED, game_offset + g), range(n_games)))
    return time.perf_counter() - t0


def _reference_available() -> bool:
    from oracle import make_ref

    return make_ref.available()


REF_SOURCE = ("oracle/_ref: the reference's own Python engine (engine/game.py BlokusGame(enable_telemetry=False), default "
              "frontier+bitboard generator, uniform random.Random choice), byte-compiled by oracle/make_ref.py")


def run_reference(args) -> None:
    """CPU arm.  With oracle/_ref present (built in the build container, ships with the tree) this times the PYTHON
    reference itself, one process per host core; otherwise the C port of it (oracle/blokus_oracle.c) on all threads."""
    rank = _env_int("RANK", 0)
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    extra = {}
    use_ref = _reference_available()
    pool = None
    if use_ref:
        from oracle import ref_bench

        try:
            pool = ref_bench.PlayoutPool(cores)
        except Exception as exc:                      # report and time the port instead
            extra["reference_error"] = repr(exc)[:200]
            use_ref = False
    if use_ref:
        kind, per_worker = "reference", 2
        sample = cores * per_worker
        for _ in range(args.warmup):
            pool.step(1)
        t = 0.0
        for _ in range(args.steps):
            t += pool.step(per_worker)[1]
        pool.close()
        what = (f"{sample} games/step ({per_worker} per process, {cores} processes) x {args.steps} steps of the same "
                f"workload; {REF_SOURCE}")
        # the C port beside it, same box, all threads (what round 1 reported as the reference arm)
        tp = _cpu_playouts(max(cores * 8, 64), cores, 0)
        extra["port_playouts_per_sec"] = max(cores * 8, 64) / tp
    else:
        kind = "port"
        sample = max(cores * 16, 64)                  # games per step: a bounded sample of the workload
        for w in range(args.warmup):
            _cpu_playouts(max(cores, 8), cores, 10_000_000 + w * 1000)
        t = 0.0
        for k in range(args.steps):
            t += _cpu_playouts(sample, cores, k * sample)
        what = (f"{sample} games/step x {args.steps} steps of the same workload (oracle/blokus_oracle.c: C restatement "
                "of the reference's default frontier+bitboard engine; oracle/_ref was not built in this tree)")
    value = sample * args.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "games_per_step": sample},
        "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": what}, **extra),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def run_cuda(args) -> None:
    import torch
    import torch.distributed as dist

    from reinforcementlearning_blokus_b200 import device as dev
    from reinforcementlearning_blokus_b200 import _lib
    from reinforcementlearning_blokus_b200.state import empty_states

    rank, world, local = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback; use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.games
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2

    def step_device(k):
        # global game index = (k * world + rank) * n + i : disjoint Philox streams over steps and ranks
        dev.playout(n, seed=SEED, game_offset=(k * world + rank) * n, out=out, want_board_scores=False)

    sampler = ClockSampler(local)
    sampler.start()                                # nvidia-smi needs ~1 s to start: begin before the warm-up
    for w in range(args.warmup):
        step_device(1_000_000 + w)
    # keep the GPU under the same load until the sampler delivers, so the clocks line covers the timed region
    t_wait = time.perf_counter()
    while len(sampler.lines) < 3 and time.perf_counter() - t_wait < 3.0:
        step_device(1_500_000)
        torch.cuda.synchronize()
    sampler.lines.clear()
    launches0 = dev.launches()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                      # L2 flush between timed iterations (untimed)
        ev[k][0].record(stream)
        step_device(k)
        ev[k][1].record(stream)
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = dev.launches() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop()
    value = n * world * args.steps / (ms * 1e-3)

    # ---- e2e: host buffers through the public API (device.PlayoutPipeline): every step uploads its start states
    # from pinned host memory and downloads its scores to pinned host memory; the two pipeline slots overlap one
    # step's copies with the other's kernel.  Timed with CUDA events on the main stream, whose end event waits for
    # both pipeline streams; every step's H2D, kernel and D2H complete inside the timed region.
    host_states = torch.from_numpy(empty_states(n).view(np.uint8).reshape(n, 256).copy()).pin_memory()
    pipe = dev.PlayoutPipeline(n)
    mean_score = 0.0

    def run_e2e(steps, base):
        acc = 0.0
        for k in range(steps):
            pipe.submit(host_states, SEED, game_offset=((base + k) * world + rank) * n)
            got = pipe.collect()
            if got is not None:
                acc += float(got[0, 0])          # touch the result on the host
        for got in pipe.drain():
            acc += float(got[0, 0])
        return acc

    run_e2e(max(2, args.warmup), 2_000_000)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_e2e(args.steps, 0)
    for slot in pipe.slots:                      # the end event waits for every pipeline stream
        stream.wait_stream(slot["stream"])
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = n * world * args.steps / (ms_e2e * 1e-3)
    mean_score = float(pipe.slots[0]["hscores"].float().mean())

    # ---- root-parallel MCTS (the third part of BASELINE.json's metric): every rank searches its own trees from
    # the empty board (50-ply rollouts), then ONE all-reduce of the root statistics (dist.merge_root_stats)
    mcts = None
    try:   # the search extras must never take the headline line down with them (a failure is reported in `extras`)
        if not args.no_extras:
            from reinforcementlearning_blokus_b200 import dist as bdist
            from reinforcementlearning_blokus_b200.mcts import DeviceUCT

            trees, iters = 4 * 3552, 256        # four waves of 24 one-warp trees per SM, handed out by a ticket counter
            root = dev.to_device(empty_states(1))
            eng = DeviceUCT(iters, n_trees=trees, seed=SEED, tree_offset=rank * trees)
            peer = bdist.PeerRootMerge()                  # windows over NVLink peer memory (CUDA IPC), mapped once
            res = eng.search(root, iters)                 # warm-up (search and merge: the first reduction loads torch kernels)
            n_root0 = int(res["counts"][0].item())
            peer.merge(res["visits"], res["reward"], n_root0)
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(stream)
            res = eng.search(root, iters)
            v, w = peer.merge(res["visits"], res["reward"], n_root0, check=False)    # sums over this rank's trees, then ONE kernel
            m1.record(stream)
            barrier()
            assert int(peer.status.item()) == 0
            ms_m = m0.elapsed_time(m1)
            if world > 1:
                t = torch.tensor([ms_m], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_m = float(t.item())
            mcts = {"mcts_sims_per_sec": world * trees * iters / (ms_m * 1e-3), "mcts_trees_per_gpu": trees,
                    "mcts_iterations_per_tree": iters, "mcts_rollout_plies": 50,
                    "mcts_root_visits_total": int(v.sum().item()), "mcts_collective": "1 bk_root_merge_p2p per search (int64[2,n_root] added into every rank's window over NVLink peer memory)"}
            # the reference's default rollout policy (HeuristicAgent softmax, mcts_agent.py:275-276), no collective
            peer.close()
            del eng, peer
            ht = 2 * 148 * 20                   # two waves of 20 one-warp trees per SM (BK_MCTS_HEUR_CTAS_PER_SM)
            heng = DeviceUCT(iters, n_trees=ht, seed=SEED, tree_offset=rank * ht, rollout_policy=dev.POLICY_HEURISTIC)
            heng.search(root, iters)
            torch.cuda.synchronize()
            m0.record(stream)
            heng.search(root, iters)
            m1.record(stream)
            torch.cuda.synchronize()
            mcts["mcts_heuristic_rollout_sims_per_sec_per_gpu"] = ht * iters / (m0.elapsed_time(m1) * 1e-3)
            del heng

        # ---- STRONG scaling of ONE move: a fixed budget split over 16 root-parallel wide trees, 16 / N per GPU,
        # one all-reduce of the root statistics; the merged result is the same for every N (checked by the checksum below)
        if not args.no_extras and 16 % world == 0:
            from reinforcementlearning_blokus_b200 import dist as bdist

            mid = dev.playout(64, seed=0xB10C05, max_turns=24, want_final=True)["final_states"][17:18].contiguous()
            strong = {}
            for tag, sims, kw in (("uct_heuristic_rollouts_16384_sims", 16384, dict(rollout_policy=dev.POLICY_HEURISTIC)),
                                  ("flat_uniform_rollouts_8192_sims", 8192, dict(max_depth=1, ucb_variant=1, reward_mode=1))):
                rp = bdist.RootParallelWide(sims, 16, seed=SEED, **kw)   # merge = one kernel over NVLink peer memory
                out = rp.search(mid)                                  # warm-up (allocations, window mapping)
                rp.search(mid, search_index=1)
                n_root = out["n_root"]
                bdist.merge_root_stats(out["visits"], out["reward"])  # NCCL channel setup for the comparison below
                barrier()
                reps, t_search, t_merge, t_nccl = 5, 0.0, 0.0, 0.0
                a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
                for k in range(reps):
                    rp.eng.p.tree_offset = rp.rank * rp.trees_per_rank
                    barrier()
                    a.record(stream)
                    res = rp.eng.search(mid, rp.per_tree)
                    b.record(stream)
                    v, w = rp.peer.merge(res["visits"], res["reward"], n_root, check=False)
                    c.record(stream)
                    v2, w2 = bdist.merge_root_stats(res["visits"].sum(dim=0), res["reward"].sum(dim=0))   # the NCCL way, for comparison
                    d.record(stream)
                    torch.cuda.synchronize()
                    assert int(rp.peer.status.item()) == 0 and torch.equal(v, v2[:n_root]) and torch.equal(w, w2[:n_root])
                    t_search += a.elapsed_time(b)
                    t_merge += b.elapsed_time(c)
                    t_nccl += c.elapsed_time(d)
                tt = torch.tensor([t_search / reps, t_merge / reps, t_nccl / reps], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                strong[tag] = {"ms_per_move": float(tt[0] + tt[1]), "search_ms": float(tt[0]), "merge_us": float(tt[1]) * 1e3,
                               "merge": "bk_root_merge_p2p (one kernel, 64-bit atomics on NVLink peer memory)",
                               "merge_us_nccl_all_reduce": float(tt[2]) * 1e3,
                               "trees_total": 16, "trees_per_gpu": 16 // world, "sims_per_tree": rp.per_tree, "round_width": rp.width,
                               "best_rank": int(torch.argmax(v[:n_root]).item()),
                               "visits_checksum": int((v[:n_root] * torch.arange(1, n_root + 1, device=v.device)).sum().item())}
                rp.peer.close()
                del rp
            if mcts is not None:
                mcts["mcts_strong_scaling_one_move"] = strong
    except Exception as exc:      # noqa: BLE001 -- reported, not hidden: the line carries the message
        mcts = dict(mcts or {}, mcts_extras_error=repr(exc)[:300])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_playout_uniform; one launch per step)
    hbm_peak, peak_src, sm_max = _measured_peaks()
    per_launch_s = ms * 1e-3 / args.steps
    hbm_achieved = ALG_BYTES_PER_PLAYOUT_DEVICE * n / per_launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_achieved / hbm_peak, "traffic": _traffic_from_profiles(),
                "peak_source": peak_src, "kernel": "k_playout<0,false>",
                "note": "integer-ALU bound by design: HBM fraction is expected to be tiny; see int_roofline"}
    int_roofline = _int_roofline(n, per_launch_s, clocks, sm_max)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "games_per_gpu": n, "seed": SEED, "rng": "philox4x32-10",
                   "l2": "256 MiB buffer written between timed steps (flush)", "parallelism": f"games sharded x{world}"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 256, "d2h_bytes_per_step": n * 16,
                "ms_per_step": ms_e2e / args.steps, "mean_game_score": mean_score},
        "gpu_launches": launches,
        "roofline": roofline,
        "int_roofline": int_roofline,
        "lib": _lib.load().bk_version().decode(),
    }
    if not args.no_extras:
        line["extras"] = dict(mcts or {})
    if world == 1 and not args.no_extras:
        try:
            line["extras"].update(_extras(dev, torch, n))
        except Exception as exc:      # noqa: BLE001 -- same rule: report it in the line, keep the headline
            line["extras"]["extras_error"] = repr(exc)[:300]
        # the one HBM-bound kernel of the path: the legal-anchor bitmap generator (256 B in, 7 280 B out per state)
        gps = line["extras"].get("movegen_bitmap_gens_per_sec")
        if gps is not None:
            line["roofline_movegen_bitmap"] = {"bound": "hbm", "kernel": "k_movegen_bitmap", "achieved": gps * (256 + 7280) / 1e9,
                                               "peak": hbm_peak, "unit": "GB/s", "frac": gps * (256 + 7280) / 1e9 / hbm_peak,
                                               "peak_source": peak_src, "bytes_per_state": 256 + 7280}
        line["cpu_baseline"] = _cpu_baseline()
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _playout_constants():
    """ncu-measured constants of k_playout<0,false> (warp instructions per playout, DRAM bytes per launch) with the hash
    of the kernel sources they were measured on; `stale` when the kernel has changed since (tools/playout_traffic.py
    regenerates the file from an ncu capture)."""
    from reinforcementlearning_blokus_b200 import build

    for name in ("r02_playout_traffic.json", "r01_playout_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                with open(p) as f:
                    d = json.load(f)
            except Exception:
                continue
            d["file"] = "profiles/" + name
            d["stale"] = d.get("kernel_source_hash") != build.playout_source_hash()
            return d
    return {}


def _traffic_from_profiles():
    return _playout_constants().get("dram_bytes_per_launch")


def _int_roofline(n, per_launch_s, clocks, sm_max):
    """Issue-slot roofline: warp instructions issued per second vs 148 SMs x 4 schedulers x 1 instr/clk."""
    import ctypes as C

    res = {"bound": "int_issue", "unit": "warp-instr/s"}
    sm_mhz = clocks.get("sm_mhz") or sm_max
    peak = 148 * 4 * sm_mhz * 1e6
    res["peak"] = peak
    res["peak_source"] = f"148 SMs x 4 SMSP x 1 warp-instr/clk x {sm_mhz:.0f} MHz (median under load)"
    consts = _playout_constants()
    wi = consts.get("warp_instr_per_playout", ALG_WARP_INSTR_PER_PLAYOUT)
    res["constants_from"] = consts.get("file")
    res["stale"] = bool(consts.get("stale", True))       # True: the kernel sources changed since the ncu capture
    if wi:
        res["warp_instr_per_playout"] = wi
        res["achieved"] = wi * n / per_launch_s
        res["frac"] = res["achieved"] / peak
    # measured pipe throughput (csrc/probe.cu), warp-instr/clk/SMSP
    probe = os.path.join(ROOT, "reinforcementlearning_blokus_b200", "libbk_probe.so")
    if os.path.exists(probe):
        try:
            L = C.CDLL(probe)
            L.bk_probe_int.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
            per_trip = {0: 192, 1: 120}     # SASS instructions per loop trip (cuobjdump of libbk_probe.so)
            blocks, iters = 148 * 8, 20000
            for mode, key in ((0, "probe_alu_only_ipc"), (1, "probe_alu_fma_mix_ipc")):
                msf = C.c_float()
                if L.bk_probe_int(mode, blocks, iters, C.byref(msf)) == 0:
                    warp_instr = blocks * 8 * iters * per_trip[mode]
                    res[key] = warp_instr / (msf.value * 1e-3) / (148 * 4 * sm_mhz * 1e6)
        except Exception as e:  # measurement aid only
            res["probe_error"] = str(e)
    return res


def _extras(dev, torch, n):
    """The other two parts of BASELINE.json's metric: legal-move generations/s (and later MCTS sims/s)."""
    from reinforcementlearning_blokus_b200.state import empty_states as empty_states_fn

    def timeit(fn, reps=5):
        fn(); fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # corpus of SURVEY 8d C1-ii: 2^20 states, game g = uniform-random self-play from the empty board stopped after k
    # turns, k spread evenly over 0..60 (every stage of the game, terminal states included); 256 MiB of states, twice
    # the L2, so every generation streams its state from HBM
    parts = []
    per = -(-(1 << 20) // 61)
    for k in range(61):
        if k == 0:
            parts.append(dev.to_device(empty_states_fn(per)))
        else:
            parts.append(dev.playout(per, seed=0xB10C05, game_offset=k * per, max_turns=k, want_final=True)["final_states"].clone())
    corpus = torch.cat(parts)[: 1 << 20].contiguous()
    del parts
    m = corpus.shape[0]
    ex = {"movegen_states": m, "movegen_mean_legal_moves": float(dev.movegen_count(corpus).float().mean())}
    ms = timeit(lambda: dev.movegen_count(corpus))
    ex["movegen_count_gens_per_sec"] = m / ms * 1e3
    ms = timeit(lambda: dev.movegen_bitmap(corpus))
    ex["movegen_bitmap_gens_per_sec"] = m / ms * 1e3
    ms = timeit(lambda: dev.movegen(corpus))
    ex["movegen_csr_gens_per_sec"] = m / ms * 1e3
    # HBM view of the generators: algorithmic bytes = 256 B state in + output (4 B count | 7280 B bitmap | 2 B per id)
    hbm, _, _ = _measured_peaks()
    ex["movegen_bitmap_hbm_frac"] = ex["movegen_bitmap_gens_per_sec"] * (256 + 7280) / 1e9 / hbm
    ex["movegen_count_hbm_frac"] = ex["movegen_count_gens_per_sec"] * (256 + 4) / 1e9 / hbm
    # HeuristicAgent-policy playouts (row N1)
    hout = {}
    nh = n // 4
    ms = timeit(lambda: dev.playout(nh, seed=SEED, policy=dev.POLICY_HEURISTIC, out=hout, want_board_scores=False), reps=3)
    ex["heuristic_playouts_per_sec"] = nh / ms * 1e3
    # BASELINE config 0: the reference's micro-benchmark protocol (benchmarks/benchmark_move_generation.py:136-224) --
    # its five seeded states, side to move, mean milliseconds of single-board get_legal_moves calls through the
    # drop-in object API (one upload, one kernel, one download, Python Move objects); reference: 2.36-8.01 ms
    try:
        import json as _json

        from reinforcementlearning_blokus_b200 import Board, LegalMoveGenerator
        from reinforcementlearning_blokus_b200.state import STATE_DTYPE

        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "kat1_states.json")) as f:
            kat = _json.load(f)
        gen, per_state = LegalMoveGenerator(), []
        for k in kat:
            board = Board.from_packed(np.frombuffer(bytes.fromhex(k["packed_hex"]), dtype=STATE_DTYPE)[0])
            gen.get_legal_moves(board, board.current_player)
            t0 = time.perf_counter()
            for _ in range(50):
                gen.get_legal_moves(board, board.current_player)
            per_state.append(round((time.perf_counter() - t0) / 50 * 1e3, 4))
        ex["movegen_single_board_ms_kat1"] = per_state
    except FileNotFoundError:
        pass
    # BASELINE config 4: one agent move of ~1,000 simulations on a mid-game root (flat UCB bandit of FastMCTSAgent, 50-ply
    # rollouts): one warp per tree vs 16-way leaf-batched rollouts (63 iterations x 16 = 1,008 simulations)
    try:
        from reinforcementlearning_blokus_b200.mcts import DeviceUCT

        root = corpus[24 * per + 17: 24 * per + 18].contiguous()            # a 24-turn position
        for tag, iters, wb in (("mcts_move_1000_sims_ms_one_warp", 1000, 1), ("mcts_move_1008_sims_ms_leaf_batch16", 63, 16)):
            eng = DeviceUCT(iters, n_trees=1, max_depth=1, ucb_variant=1, reward_mode=1, seed=SEED, leaf_batch=wb)
            eng.search(root, iters)
            torch.cuda.synchronize()
            ex[tag] = round(timeit(lambda: eng.search(root, iters), reps=3), 3)
        # the same move searched by the whole GPU (K6, the round-synchronous search; what FastMCTSAgent.think runs): one tree,
        # eight root-parallel trees for the 5,000-simulation /think budget, and the full tree with HeuristicAgent rollouts
        from reinforcementlearning_blokus_b200.mcts import WideUCT

        flat = dict(max_depth=1, ucb_variant=1, reward_mode=1)
        for tag, sims, trees, kw in (("mcts_move_1000_sims_ms_wide_1tree", 1000, 1, flat), ("mcts_move_1000_sims_ms_wide_2trees", 500, 2, flat),
                                     ("mcts_move_5000_sims_ms_wide_8trees", 625, 8, flat), ("mcts_move_5000_sims_ms_wide_2trees", 2500, 2, flat),
                                     ("mcts_move_1000_sims_ms_wide_uct_heuristic_rollouts", 1000, 1, dict(rollout_policy=dev.POLICY_HEURISTIC))):
            weng = WideUCT(sims, width=max(64, min(1024, sims // 4)), n_trees=trees, seed=SEED, **kw)
            weng.search(root, sims)
            ex[tag] = round(timeit(lambda: weng.search(root, sims), reps=5), 3)
            del weng
        # UCT with learned leaf evaluation in the kernel (reward_mode 3: a pairwise logistic win-probability model on the
        # 34 snapshot features of every new leaf instead of a rollout); synthetic weights, trees from the empty board
        wts = torch.from_numpy(np.random.default_rng(0).normal(0.0, 0.05, 35))
        lt = 4 * 2368
        leng = DeviceUCT(256, n_trees=lt, seed=SEED, reward_mode=3, leaf_model=wts)
        start = dev.to_device(empty_states_fn(1))
        leng.search(start, 256)
        torch.cuda.synchronize()
        ex["mcts_learned_leaf_sims_per_sec"] = lt * 256 / timeit(lambda: leng.search(start, 256), reps=2) * 1e3
        del leng
    except Exception as exc:                                                # extras never break the headline line
        ex["mcts_move_error"] = str(exc)[:200]
    # win-probability snapshot rows (row N4): 4 x 34 float64 per board = four piece-count generations + a flood fill;
    # and a whole dataset pass: replay recorded games, capture 8 checkpoints each
    sub = corpus[::4].contiguous()                      # every fourth state: all game stages
    ms = timeit(lambda: dev.snapshot_features(sub), reps=3)
    ex["snapshot_boards_per_sec"] = sub.shape[0] / ms * 1e3
    tr = dev.playout(n, seed=SEED, want_trace=True)["trace"]
    ms = timeit(lambda: dev.replay_snapshots(tr, (8, 16, 24, 32, 40, 48, 56, 64)), reps=3)
    ex["snapshot_replay_games_per_sec"] = n / ms * 1e3
    return ex


def _cpu_baseline():
    """The reference's Python engine on all host cores (oracle/_ref, bounded sample) with its movegen benchmark and
    search agents on one core; beside it the C port of the engine (oracle/blokus_oracle.c) on one core."""
    out = {}
    try:
        if _reference_available():
            out = _cpu_reference_baseline()
    except Exception as exc:
        out = {"reference_error": repr(exc)[:200]}
    out["port"] = _cpu_port_baseline()
    if not out.get("kind"):
        out.update({k: out["port"][k] for k in ("value", "unit", "cores", "kind", "sample")})
    return out


def _cpu_reference_baseline():
    """oracle/_ref on all host cores (playouts), and on one core: its movegen benchmark and its search agents."""
    from oracle import ref_bench

    cores = os.cpu_count() or 1
    budget = float(os.environ.get("BLOKUS_CPU_BASELINE_SECONDS", "12"))
    pool = ref_bench.PlayoutPool(cores)
    pool.step(1)
    games, t = 0, 0.0
    while t < budget:
        g, dt = pool.step(2)
        games, t = games + g, t + dt
    pool.close()
    out = {"value": games / t, "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": f"{games} games of the same workload in {t:.1f} s on {cores} processes; {REF_SOURCE}",
           "playouts_per_sec_per_core": games / t / cores}
    # BASELINE configs[0]: benchmarks/benchmark_move_generation.py's protocol, one core
    mg = ref_bench.movegen_benchmark(num_runs=10)
    out["movegen_frontier_bitboard_gens_per_sec_one_core"] = mg["frontier_bitboard_gens_per_sec"]
    out["movegen_naive_gens_per_sec_one_core"] = mg["naive_gens_per_sec"]
    out["movegen_benchmark_ms"] = [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()} for r in mg["states"]]
    out.update({k: round(v, 3) for k, v in ref_bench.search_rates(4.0).items()})
    return out


def _cpu_port_baseline():
    """Oracle port on ONE host core, bounded sample (~8 s)."""
    budget = float(os.environ.get("BLOKUS_CPU_PORT_SECONDS", "8"))
    from oracle import orc
    orc.lib()
    t0 = time.perf_counter()
    g = 0
    while time.perf_counter() - t0 < budget:
        orc.playout(SEED, g)
        g += 1
    dt = time.perf_counter() - t0
    out = {"value": g / dt, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{g} games of the same workload in {dt:.1f} s (oracle/blokus_oracle.c, default "
                     "frontier+bitboard generator, game-over check after every move as engine/game.py:182)"}
    # the other two parts of the metric on the same core: legal-move generations/s on mid-game states and UCT
    # simulations/s with the bench's search settings (uniform 50-ply rollouts, transposition table off)
    mids = [orc.playout(SEED, k, max_turns=6 * (1 + k % 8), want_final=True)["final"] for k in range(64)]
    boards = [orc.OracleBoard.from_packed(m) for m in mids]
    t0 = time.perf_counter(); gens = 0
    while time.perf_counter() - t0 < 2.0:
        for b in boards:
            b.legal_moves(b.current_player)
        gens += len(boards)
    out["movegen_gens_per_sec"] = gens / (time.perf_counter() - t0)
    root = np.zeros(1, dtype=orc.STATE_DTYPE); root["first_mask"] = 0xF
    t0 = time.perf_counter(); sims = 0
    while time.perf_counter() - t0 < 3.0:
        orc.mcts_search(root, 256, SEED, tree_id=sims // 256, rollout_policy=0)
        sims += 256
    out["mcts_sims_per_sec"] = sims / (time.perf_counter() - t0)
    t0 = time.perf_counter(); sims = 0
    while time.perf_counter() - t0 < 3.0:                       # the reference's default: HeuristicAgent rollouts
        orc.mcts_search(root, 64, SEED, tree_id=sims // 64, rollout_policy=1)
        sims += 64
    out["mcts_heuristic_rollout_sims_per_sec"] = sims / (time.perf_counter() - t0)
    return out


def _emit(line: dict) -> None:
    """The ONE line of stdout: everything else that libraries print (e.g. NCCL's version banner when NCCL_DEBUG is
    set) was sent to stderr by main()."""
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


_STDOUT_FD = 1


def main():
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)          # keep the real stdout for the JSON line ...
    os.dup2(2, 1)                   # ... and route everything else written to fd 1 (C libraries included) to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
