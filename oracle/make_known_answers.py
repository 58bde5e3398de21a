"""Extract the known-answer vectors the reference's own simulator tests hold for the TJM path into
tests/golden/known_answers_simulator.json.

The vectors are the ``expected_*`` arrays embedded in /root/reference/tests/test_simulator.py (two-site correlators XX / YY / ZZ of a
noise-free Ising chain through Simulator.run: analog at the left boundary, the centre and the right boundary, and the Trotter circuit
through the digital path).  They are read with ``ast`` (nothing of the reference is executed or copied as code); the set-ups live in
tests/test_reference_known_answers.py.  Run here (the reference is not on the GPU box):  python oracle/make_known_answers.py
"""
from __future__ import annotations

import ast
import json
import os

REF_TEST = "/root/reference/tests/test_simulator.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "known_answers_simulator.json")
WANTED = ["test_two_site_correlator_left_boundary", "test_two_site_correlator_center", "test_two_site_correlator_right_boundary",
          "test_two_site_correlator_center_circuit"]


def main() -> None:
    tree = ast.parse(open(REF_TEST).read())
    out: dict = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            vecs = {}
            for stmt in ast.walk(node):
                if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name) \
                        and stmt.targets[0].id.startswith("expected_") and isinstance(stmt.value, ast.Call) and stmt.value.args:
                    vecs[stmt.targets[0].id[len("expected_"):]] = [float(x) for x in ast.literal_eval(stmt.value.args[0])]
            out[node.name] = {"source": f"tests/test_simulator.py:{node.lineno}", "expected": vecs}
    missing = [w for w in WANTED if w not in out]
    if missing:
        raise SystemExit(f"not found in the reference's tests: {missing}")
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    print(OUT, {k: {n: len(v) for n, v in d["expected"].items()} for k, d in out.items()})


if __name__ == "__main__":
    main()
