"""The reference's standard arena experiment (scripts/arena_config.json: four gameplay_fast_mcts agents with 25 / 50 /
100 / 200 ms budgets, 20 iterations per ms under its deterministic budget rule, 100 games, round robin, snapshots at
the default plies) run as device batches; prints games/s and the per-agent results."""
import json
import sys
import time
import torch
sys.path.insert(0, ".")
from reinforcementlearning_blokus_b200 import arena

cfg = arena.RunConfig.from_dict({
    "agents": [{"name": f"mcts_{ms}ms", "type": "gameplay_fast_mcts", "thinking_time_ms": ms,
                "params": {"deterministic_time_budget": True, "iterations_per_ms": 20.0}} for ms in (25, 50, 100, 200)],
    "num_games": int(sys.argv[1]) if len(sys.argv) > 1 else 100, "seed": 20260301, "seat_policy": "round_robin",
    "max_turns": 2500, "snapshots": {"enabled": True}})
arena.run_device_arena_agents(arena.RunConfig.from_dict(dict(cfg.to_dict(), num_games=4)))        # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
out = arena.run_device_arena_agents(cfg)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
sims = sum(s["total_simulations"] for s in out["agent_move_stats"].values())
print(json.dumps({"games": cfg.num_games, "seconds": round(dt, 2), "games_per_sec": round(cfg.num_games / dt, 2),
                  "simulations": int(sims), "sims_per_sec": round(sims / dt), "snapshot_rows": len(out["snapshot_rows"]),
                  "agents": out["summary"]["agents"]}))
