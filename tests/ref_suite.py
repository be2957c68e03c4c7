"""Runs the reference's OWN test modules (byte-compiled into oracle/_ref by oracle/make_ref.py) in a child pytest.

The modules are source-less ``.pyc`` files, which pytest does not collect; each one gets a one-line stub
(``from tests.<module> import *``) in a scratch directory, and a scratch ``conftest.py`` puts the reference tree on
``sys.path`` and -- when asked -- installs this package's device generator and device agents into the reference
(``integration.install(replace_class=True, patch_board=True)`` + ``install_agents()``) before anything is collected.
"""

from __future__ import annotations

import os
import re
import subprocess
import sys
from typing import Dict, Sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFTEST = '''import os, sys
sys.dont_write_bytecode = True
sys.path.insert(0, {tree!r})
sys.path.insert(1, {root!r})
DEVICE = {device!r}
if DEVICE:
    import engine.move_generator, engine.game, agents.fast_mcts_agent, mcts.mcts_agent   # load before patching
    import analytics.tournament.arena_runner
    from reinforcementlearning_blokus_b200 import integration
    integration.install(replace_class=True, patch_board=True)
    integration.install_agents()
    # one warm call: CUDA context + the library (the reference's latency test allows 700 ms for a 150 ms budget)
    from engine.game import BlokusGame
    from agents.fast_mcts_agent import FastMCTSAgent
    _g = BlokusGame()
    _p = _g.get_current_player()
    FastMCTSAgent(iterations=64, time_limit=0.05).think(_g.board, _p, _g.get_legal_moves(_p), 20)


import engine.move_generator as _mg
import engine.board as _eb
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "install_report.txt"), "w") as _f:
    _f.write("generator class: %s; shared: %s; FastMCTSAgent: %s; MCTSAgent: %s; Board.can_place_piece: %s" % (
        _mg.LegalMoveGenerator.__name__, type(_mg.get_shared_generator()).__name__,
        __import__("agents.fast_mcts_agent").fast_mcts_agent.FastMCTSAgent.__name__,
        __import__("mcts.mcts_agent").mcts_agent.MCTSAgent.__name__, _eb.Board.can_place_piece.__name__))
'''


def run(modules: Sequence[str], scratch: str, device: bool, timeout: int = 1500, extra_args: Sequence[str] = ()) -> Dict:
    """modules: dotted names like ``tests.test_engine``.  Returns dict(returncode, passed, failed, errors, output)."""
    sys.path.insert(0, ROOT)
    from oracle import make_ref

    tree = make_ref.ref_path()
    os.makedirs(scratch, exist_ok=True)
    with open(os.path.join(scratch, "conftest.py"), "w") as f:
        f.write(CONFTEST.format(tree=tree, root=ROOT, device=bool(device)))
    files = []
    for m in modules:
        name = "test_ref_" + m.split(".")[-1].removeprefix("test_") + ".py"
        with open(os.path.join(scratch, name), "w") as f:
            f.write(f"from {m} import *  # noqa: F401,F403 -- the reference's own tests, byte-compiled in oracle/_ref\n")
        files.append(name)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONPATH", None)
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", scratch, "-c", os.devnull,
           "-o", "python_files=test_ref_*.py", *extra_args, *files]
    proc = subprocess.run(cmd, cwd=scratch, env=env, capture_output=True, text=True, timeout=timeout)
    out = proc.stdout + proc.stderr
    counts = {k: 0 for k in ("passed", "failed", "error", "skipped")}
    for n, k in re.findall(r"(\d+) (passed|failed|errors?|skipped)", out.strip().splitlines()[-1] if out.strip() else ""):
        counts[k.rstrip("s") if k.startswith("error") else k] = int(n)
    report = ""
    try:
        with open(os.path.join(scratch, "install_report.txt")) as f:
            report = f.read().strip()
    except OSError:
        pass
    return {"returncode": proc.returncode, "output": out, "report": report, **counts}
