#!/usr/bin/env python
"""bench.py - self-play search throughput (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch: a batched MCTS.run (root inference +
num_simulations x {select, recurrent inference, expand, backup}) for every game of the batch,
i.e. one env-step's worth of search per game.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

N=1 workload: BASELINE.json configs[1] - CartPole, fully-connected net, num_simulations=50,
4096 parallel games per GPU (weak scaling: every rank owns its own 4096 games; no data-path
collective - only one NCCL all-gather of per-rank counters per reporting step).

Prints ONE JSON line (see the keys at the bottom).  `value` = env-steps/s with the inputs
resident in HBM; `e2e` = the same through the C ABI with pinned HOST buffers (H2D + D2H inside
the timed region); `roofline` = algorithmic tree+hidden bytes of the dominant kernel over its
CUDA-event duration against the measured HBM peak; `cpu_baseline` = the oracle port of the
reference's batch-1 Python/torch search timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (game, games per GPU, num_simulations, algorithmic tree+hidden bytes per simulation)
    "cartpole_b4096_n50": ("cartpole", 4096, 50, 633.0),
    "cartpole_b4096_n25": ("cartpole", 4096, 25, 530.0),
    "tictactoe_b8192_n50": ("tictactoe", 8192, 50, 2420.0),
    "connect4_b1024_n200": ("connect4", 1024, 200, 22300.0),
    "breakout_b128_n50": ("breakout", 128, 50, 5100.0),
}
DEFAULT_WORKLOAD = "cartpole_b4096_n50"


def load_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s, which) from the driver-written MEASURED_PEAKS.json, else the fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


# algorithmic FLOPs of initial_inference / recurrent_inference per sample
NET_FLOPS = {"cartpole": (1312.0, 2752.0), "tictactoe": (1.880e5, 2.315e5), "connect4": (3.737e7, 4.040e7),
             "breakout": (3.419e7, 1.532e6)}


# measured DRAM bytes per launch of the dominant kernel (ncu --set full, profiles/): fc_search / conv tower
TRAFFIC = {"connect4": 8919424}      # conv_tower_resident_kernel: (8966144 + 8872704) / 2 bytes read, 0 written back within the capture


def conv3x3_flops(spec, N):
    """FLOPs of the 3x3 convolutions of one search per game: initial_inference + N recurrent_inferences (models.py)."""
    C, blocks = spec.channels, spec.blocks
    obs_c, (_, H, W) = spec.in_channels, spec.obs_shape
    total = 0.0
    if spec.downsample:
        conv = lambda h, w, ci, co: 2.0 * h * w * ci * co * 9
        h1, w1 = (H + 1) // 2, (W + 1) // 2
        total += conv(h1, w1, obs_c, C // 2) + 2 * 2 * conv(h1, w1, C // 2, C // 2)
        h2, w2 = (h1 + 1) // 2, (w1 + 1) // 2
        total += conv(h2, w2, C // 2, C) + 3 * 2 * conv(h2, w2, C, C)
        h3, w3 = (h2 + 1) // 2, (w2 + 1) // 2
        total += 3 * 2 * conv(h3, w3, C, C)
        H, W = (h3 + 1) // 2, (w3 + 1) // 2
    else:
        total += 2.0 * H * W * obs_c * C * 9
    block = 2 * 2.0 * H * W * C * C * 9
    total += 2 * blocks * block                                        # representation + prediction towers
    total += N * (2.0 * H * W * (C + 1) * C * 9 + 2 * blocks * block)  # dynamics stem + dynamics / prediction towers
    return total


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(numpy.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    game, n_sim, seconds, seed = args
    import torch
    torch.set_num_threads(1)
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config, synthetic_weights
    from oracle import mcts as om
    from oracle.net import OracleNet
    cfg = load_game_module(game).MuZeroConfig()
    spec = netspec_from_config(cfg)
    net = OracleNet(spec, synthetic_weights(spec, 0))
    params = om.SearchParams.from_config(cfg, n_sim)
    rs = numpy.random.RandomState(seed)
    draws = om.LegacyNumpyDraws(rs)
    search = om.TreeSearch(params)
    ev = om.ModelEvaluator(net, spec.support_size)
    shape = (spec.in_channels,) + tuple(spec.obs_shape[1:])
    legal = list(range(spec.action_space))

    def one():
        if game == "cartpole":
            obs = rs.uniform(-0.05, 0.05, size=shape).astype(numpy.float32)
        else:
            obs = rs.random_sample(shape).astype(numpy.float32)
        search.run(ev, obs, legal, 0, True, draws)

    one()                      # warm-up
    t0 = time.perf_counter()
    done = 0
    while time.perf_counter() - t0 < seconds:
        one()
        done += 1
    return done, time.perf_counter() - t0


def cpu_baseline(game, n_sim, seconds, cores):
    """env-steps/s of the oracle port (batch-1 Python/torch MCTS.run) on `cores` host processes."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(game, n_sim, seconds, 1000 + i) for i in range(cores)])
    searches = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return searches / wall, searches, wall


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def selfplay_loop(game, B, N, device, moves):
    """env-steps/s of the full loop: BatchedSelfPlay over B games for ~`moves` lockstep moves."""
    from muzero_general_b200 import self_play as sp
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config, synthetic_weights
    mod = load_game_module(game)
    cfg = mod.MuZeroConfig()
    cfg.num_parallel_games, cfg.rng_mode, cfg.num_simulations = B, "philox", N
    spec = netspec_from_config(cfg)
    worker = sp.SelfPlay({"weights": synthetic_weights(spec, 0)}, mod.Game, cfg, seed=0, device=device)
    worker.play_games(1, 1.0, max_total_moves=2 * B)                  # warm-up (two moves)
    start_steps, start_games = worker.played_steps, worker.played_games
    t0 = time.perf_counter()
    worker.play_games(10 ** 9, 1.0, max_total_moves=start_steps + moves * B)
    dt = time.perf_counter() - t0
    steps = worker.played_steps - start_steps
    res = {"value": steps / dt, "unit": "env-steps/s", "env_steps": int(steps), "seconds": dt,
           "games_finished": int(worker.played_games - start_games),
           "includes": "mz_search + numpy vector env step + Dirichlet draw + action sampling + GameHistory assembly"}
    worker.model.engine.close()
    return res


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    game, B, N, bytes_per_sim = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": args.workload, "game": game, "games_per_gpu": B, "num_simulations": N,
              "net": "fullyconnected" if game == "cartpole" else "resnet", "weights": "synthetic seed 0",
              "l2": "256 MiB buffer written between timed steps", "parallelism": f"games sharded x{world}"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = host_cores()
        steps, total, wall = max(1, args.steps), 0, 0.0
        per_step = max(2.0, min(20.0, 120.0 / (steps + args.warmup)))
        for _ in range(args.warmup):
            cpu_baseline(game, N, 1.0, cores)
        for _ in range(steps):
            _, s, w = cpu_baseline(game, N, per_step, cores)
            total += s; wall += w
        v = total / wall
        sample = f"{steps} steps x {per_step:.1f}s of batch-1 MCTS.run (N={N}) on {cores} processes"
        print(json.dumps({
            "impl": "reference", "metric": "self-play env-steps/sec", "value": v, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic", "config": config, "sims_per_sec": v * N,
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    from muzero_general_b200.engine import SearchEngine
    from muzero_general_b200.games import load_game_module
    from muzero_general_b200.netspec import netspec_from_config, synthetic_weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    cfg = load_game_module(game).MuZeroConfig()
    spec = netspec_from_config(cfg)
    A = spec.action_space
    eng = SearchEngine(cfg, max_games=B, device=local_rank, num_simulations=N, seed=cfg.seed + rank)
    eng.load_weights(synthetic_weights(spec, 0))

    # synthetic inputs, a different batch every step (global game ids keep streams rank-independent)
    n_batches = 4
    rs = numpy.random.RandomState(100 + rank)
    shape = (B, eng.obs_elems)
    if game == "cartpole":
        host_obs = [rs.uniform(-0.05, 0.05, size=shape).astype(numpy.float32) for _ in range(n_batches)]
    else:
        host_obs = [rs.random_sample(shape).astype(numpy.float32) for _ in range(n_batches)]
    host_noise = [rs.dirichlet([cfg.root_dirichlet_alpha] * A, size=B) for _ in range(n_batches)]
    pin = lambda a: torch.from_numpy(a).pin_memory()
    pinned_obs = [pin(a) for a in host_obs]
    pinned_noise = [pin(a) for a in host_noise]
    dev_obs = [t.to(dev) for t in pinned_obs]
    dev_noise = [t.to(dev) for t in pinned_noise]
    game_id = (rank * B + numpy.arange(B)).astype(numpy.int64)
    dev_gid = torch.from_numpy(game_id).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        return eng.search(obs=dev_obs[i % n_batches], add_exploration_noise=True, noise=dev_noise[i % n_batches],
                          game_id=dev_gid)

    def step_host(i):
        return eng.search(obs=pinned_obs[i % n_batches].numpy(), add_exploration_noise=True,
                          noise=pinned_noise[i % n_batches].numpy(), game_id=game_id)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        wall, kern, visits = 0.0, 0.0, None
        for i in range(steps):
            flush.fill_(i & 0xFF)                      # evict L2 between timed iterations (untimed)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(i)                                # mz_search synchronises its stream before returning
            torch.cuda.synchronize()
            wall += time.perf_counter() - t0
            kern += out.device_ms
            visits = out.visit_counts
        return wall, kern, visits

    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.launch_count
    wall, kern_ms, visits = timed(step_device, args.steps, args.warmup)
    launches = eng.launch_count - launches0
    clk = clocks.stop()
    wall_e2e, _, visits_h = timed(step_host, args.steps, args.warmup)
    assert int(numpy.asarray(visits_h).sum()) == B * N
    kernel_split = {}
    if game != "cartpole" and rank == 0:
        eng.kernel_timing(True)
        eng.kernel_times()
        flush.fill_(7)
        torch.cuda.synchronize()
        step_device(0)
        kernel_split = eng.kernel_times()
        eng.kernel_timing(False)

    # max over ranks + the single counter all-gather of the reporting step
    t = torch.tensor([wall, wall_e2e, kern_ms], dtype=torch.float64, device=dev)
    counts = torch.tensor([B * args.steps, B * args.steps * N], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(gathered, counts)
        total_steps = int(sum(int(g[0]) for g in gathered))
    else:
        total_steps = int(counts[0])
    wall, wall_e2e, kern_ms = (float(x) for x in t.tolist())

    if rank == 0:
        value = total_steps / wall
        hbm_peak, bf16_peak, peak_kind = load_peaks()
        kern_s = kern_ms / 1000.0 / args.steps
        if game == "cartpole":
            # dominant kernel: the fused search kernel, one launch per step (HBM roofline)
            alg_bytes = B * (N * bytes_per_sim + eng.obs_elems * 4 + A * 8 + A * 4 + 8)
            achieved = alg_bytes / kern_s / 1e9
            roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": 242944, "peak_kind": peak_kind,
                        "kernel": "fc_search_kernel", "algorithmic_bytes_per_launch": alg_bytes,
                        "note": "tree + hidden states live in shared memory for FC nets: measured DRAM traffic "
                                "(profiles/r01_fc_search_ncu.md) is 0.24 MB per launch, the kernel is issue/latency-"
                                "bound (47.6 % of peak issue rate), not HBM-bound"}
        else:
            # residual nets: tensor roofline for the dominant kernel, timed live with CUDA event pairs
            # around every launch of one extra (untimed) step (mz_kernel_timing); algorithmic FLOPs = the 3x3
            # convolutions that kernel class executes in one step (2*H*W*Cin*Cout*9 each).  fp16 tensor-core rate
            # = the measured dense bf16 figure.  The whole-step figure (all kernels) is kept as `step_level`.
            f0, f1 = NET_FLOPS[game]
            flops = B * (f0 + N * f1)
            conv_flops = B * conv3x3_flops(spec, N)
            split = {k: {"ms": v[0], "launches": v[1]} for k, v in kernel_split.items() if v[1]}
            total_ms = sum(v["ms"] for v in split.values()) or 1.0
            for v in split.values():
                v["share"] = v["ms"] / total_ms
            dominant = max(("conv_tower_tc_kernel", "conv3x3_kernel"), key=lambda k: split.get(k, {"ms": 0.0})["ms"])
            dom = split.get(dominant, {"ms": total_ms, "launches": 1})
            achieved = conv_flops / (dom["ms"] / 1000.0) / 1e12
            roofline = {"bound": "tensor", "achieved": achieved, "peak": bf16_peak, "unit": "TFLOP/s",
                        "frac": achieved / bf16_peak, "traffic": TRAFFIC.get(game), "peak_kind": peak_kind + " dense bf16 (sustained)",
                        "kernel": dominant, "launches_per_step": dom["launches"],
                        "avg_launch_us": 1000.0 * dom["ms"] / max(dom["launches"], 1),
                        "algorithmic_flops_per_launch": conv_flops / max(dom["launches"], 1),
                        "kernel_split": split,
                        "step_level": {"algorithmic_flops_per_step": flops, "achieved": flops / kern_s / 1e12,
                                       "frac": flops / kern_s / 1e12 / bf16_peak},
                        "note": ("tcgen05 towers: conv_tower_resident_kernel up to 1184 boards per launch (activations stay in "
                                 "shared memory), conv_tower_tc_kernel above; fp16 operands, fp32 accumulate; 84 of 128 rows "
                                 "of every MMA are real board positions" if dominant == "conv_tower_tc_kernel" else
                                 "fp32 CUDA-core direct convolution (strict numerics): the tensor peak is not reachable by design")}
        out = {
            "metric": "self-play env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 nets + f64 tree statistics" if game != "connect4" or os.environ.get("MZ_NO_TC") == "1" else
                      "fp16 operands / f32 accumulate (tensor-core towers), f32 heads, f64 tree statistics"),
            "data": "synthetic", "config": config,
            "sims_per_sec": value * N,
            "kernel_ms_per_step": kern_ms / args.steps,
            "e2e": {"value": total_steps / wall_e2e, "unit": "env-steps/s",
                    "h2d_bytes_per_step": int(B * (eng.obs_elems * 4 + A * 8 + 8)),
                    "d2h_bytes_per_step": int(B * (A * 4 + 8 + 4 + 4 + 4 + A * 8 + 16))},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline,
        }
        if world == 1 and game in ("cartpole", "tictactoe", "connect4"):
            # the whole self-play loop through the reference-shaped API (SelfPlay.play_games): search + host
            # environment stepping (vectorised numpy envs) + action sampling + GameHistory assembly
            try:
                out["selfplay_loop"] = selfplay_loop(game, B, N, local_rank, moves=12 if game == "cartpole" else 6)
            except Exception as e:                       # never lose the headline line over the extra
                out["selfplay_loop"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            v, searches, w = cpu_baseline(game, N, args.cpu_seconds, cores)
            out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                   "sample": f"{searches} batch-1 MCTS.run calls (N={N}) in {w:.1f}s on {cores} processes"}
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
