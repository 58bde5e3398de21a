"""All reference-side walks in one process (one import of torch and one load of the reference through the shim): used by
tests/test_reference_walks.py with a few draws each, by hand with many.  TEST INFRASTRUCTURE, build container only.

    python -m oracle.run_reference_walks --cases 10 --seed 5
"""
from __future__ import annotations

import argparse
import sys


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=10)
    ap.add_argument("--seed", type=int, default=5)
    args = ap.parse_args()
    from oracle import fuzz_reference_digital_simulator, fuzz_reference_program_simulator, fuzz_reference_simulator, fuzz_reference_vs_oracle
    walks = [("analog, reference against oracle", fuzz_reference_vs_oracle, []),
             ("digital, reference against oracle", fuzz_reference_vs_oracle, ["--digital"]),
             ("gate lists, reference against oracle", fuzz_reference_vs_oracle, ["--gates"]),
             ("programs, reference executor against oracle", fuzz_reference_vs_oracle, ["--program"]),
             ("Simulator.run (analog), reference against yaqs_b200", fuzz_reference_simulator, []),
             ("Simulator.run (programs), reference against yaqs_b200", fuzz_reference_program_simulator, []),
             ("digital_tjm worker against Simulator.run (circuits)", fuzz_reference_digital_simulator, [])]
    failed = []
    for k, (name, module, extra) in enumerate(walks):
        print(f"== {name}", flush=True)
        sys.argv = [module.__name__, "--cases", str(args.cases), "--seed", str(args.seed + k), "--keep-going", *extra]
        if module.main() != 0:
            failed.append(name)
    print("ALL WALKS CLEAN" if not failed else f"WALKS WITH MISMATCHES: {failed}")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
