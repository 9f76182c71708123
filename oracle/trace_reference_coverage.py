"""ORACLE / TEST INFRASTRUCTURE ONLY — which statements of the reference's env code do the golden trajectories execute?

    python oracle/trace_reference_coverage.py            (build container: needs /root/reference; ~6 minutes)

Runs oracle/make_golden.py's main() under the stdlib tracer (the fixtures go to a scratch directory, tests/golden is not
touched) and lists, per file of Louvre_Evacuation/envs/, the statements that were never executed.  Used to aim goldens at
cold branches (traj_room_timelimit, traj_topexit, traj_westexit_far); what is left after them is unreachable code
(evacuation_env.py:187,218,265, people.py:79), helpers off the hot path (create_risky_initial_positions, Random_Valid_Point,
the automatic robot sweep, get_performance_metrics, printing) and constructor variants of the fire model."""
import ast
import collections
import os
import sys
import tempfile
import trace

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True


def main():
    import make_golden as mg
    mg.OUT = tempfile.mkdtemp(prefix="golden_cov_")
    tr = trace.Trace(count=1, trace=0, ignoredirs=[sys.prefix, sys.exec_prefix])
    tr.runfunc(mg.main)
    executed = collections.defaultdict(set)
    for (fn, ln), _n in tr.results().counts.items():
        if "Louvre_Evacuation" + os.sep + "envs" in fn:
            executed[fn].add(ln)
    for fn in sorted(executed):
        src = open(fn, encoding="utf-8").read()
        lines = src.split("\n")
        stmts = {n.lineno for n in ast.walk(ast.parse(src))
                 if isinstance(n, ast.stmt) and not isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.Import, ast.ImportFrom))}
        print(f"===== {fn.split('Louvre_Evacuation' + os.sep)[-1]}: executed {len(stmts & executed[fn])} of {len(stmts)} statements")
        for ln in sorted(stmts - executed[fn]):
            text = lines[ln - 1].strip()
            if text and not text.startswith(('"""', "'''", "#")):
                print(f"   {ln}: {text[:110]}")


if __name__ == "__main__":
    main()
