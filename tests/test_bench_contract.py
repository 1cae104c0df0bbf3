"""Static checks of bench.py that need no GPU: the multi-rank launch must not hang.

Every rank of `torchrun bench.py --gpus N` has to enter the same collectives in the same order.  `hot_step`, `e2e_step`
and everything built on them contain the gradient all-reduce, and `barrier` is one: none of them may sit under a branch
that only some ranks take.  (Round 2 shipped such a branch for one commit; on two GPUs it ended in NCCL's watchdog.)
"""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLLECTIVE_CALLS = {"hot_step", "e2e_step", "e2e_run", "timed", "barrier", "all_reduce", "broadcast", "all_gather"}


def _mentions_rank(node):
    return any(isinstance(n, ast.Name) and n.id in ("rank", "local_rank") for n in ast.walk(node))


def _calls(node):
    for n in ast.walk(node):
        if isinstance(n, ast.Call):
            f = n.func
            yield f.id if isinstance(f, ast.Name) else (f.attr if isinstance(f, ast.Attribute) else "")
        if isinstance(n, ast.Name) and n.id in COLLECTIVE_CALLS:      # passed as a callable, e.g. timed(hot_step, n)
            yield n.id


def test_no_collective_under_a_rank_dependent_branch():
    with open(os.path.join(ROOT, "bench.py")) as fh:
        tree = ast.parse(fh.read())
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    bad = []
    for node in ast.walk(main):
        if isinstance(node, ast.If) and _mentions_rank(node.test):
            for stmt in node.body + node.orelse:
                hit = sorted(set(_calls(stmt)) & COLLECTIVE_CALLS)
                if hit:
                    bad.append((node.lineno, hit))
    assert not bad, f"collectives under a rank-dependent branch in bench.py: {bad}"


def test_bench_declares_the_contract_keys():
    """The JSON line's keys the driver reads are all spelled in bench.py (cheap guard against a rename)."""
    with open(os.path.join(ROOT, "bench.py")) as fh:
        src = fh.read()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "h2d_bytes_per_step", "d2h_bytes_per_step", "gpu_launches",
                "roofline", "cpu_baseline", "clocks", "impl"):
        assert f'"{key}"' in src, key
