"""Provenance of tests/golden/reference_known_answers.json.

The reference (Rust crate `gad` + the third-party `arrayfire` crate) cannot be built or run in this
image (no rustc/cargo, no libaf), so its outputs cannot be regenerated. The fixture holds the values
the reference's OWN tests assert, transcribed by hand with the file:line of each assertion. This
script re-derives every number from the closed form the reference test states (e.g.
`b.data() * 2.0 * direction`) so a typo in the transcription is caught:

    python tests/golden/make_golden.py          # verifies the committed file, exit code 0 if consistent
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
D = 1.7  # `let direction = 1.7;` in gad/tests/analytic.rs

EXPECTED = {
    "exp": (math.exp(2.0), [math.exp(2.0) * 2.0 * D]),
    "log": (math.log(2.0), [1.0 * D]),
    "log1p": (math.log(2.0), [0.5 * D]),
    "sin": (math.sin(1.0), [math.cos(1.0) * D]),
    "cos": (math.cos(1.0), [-math.sin(1.0) * D]),
    "tanh": (math.tanh(1.0), [(1.0 - math.tanh(1.0) ** 2) * D]),
    "sigmoid": (1.0 / (1.0 + math.exp(-1.0)), [(1.0 - 1.0 / (1.0 + math.exp(-1.0))) * (1.0 / (1.0 + math.exp(-1.0))) * D]),
    "reciprocal": (0.5, [-0.25 * D]),
    "sqrt": (math.sqrt(2.0), [0.5 / math.sqrt(2.0) * D]),
    "div": (2.0 / 3.0, [1.0 / 3.0 * D, -2.0 / 9.0 * D]),
    "pow": (8.0, [12.0 * D, math.log(2.0) * 8.0 * D]),
}


def main() -> int:
    with open(os.path.join(HERE, "reference_known_answers.json")) as f:
        golden = json.load(f)
    bad = 0
    for case in golden["scalar_f32_direction_1p7"]:
        value, grads = EXPECTED[case["name"]]
        if abs(case["value"] - value) > 1e-9 or any(abs(a - b) > 1e-9 for a, b in zip(case["grads"], grads)):
            print("MISMATCH", case["name"], case["value"], value, case["grads"], grads)
            bad += 1
    x0, x1 = golden["rosenbrock_f64"]["n2"]["x"]
    f = 100.0 * (x1 - x0 * x0) ** 2 + (1.0 - x0) ** 2
    g0 = -400.0 * x0 * (x1 - x0 * x0) - 2.0 * (1.0 - x0)
    g1 = 200.0 * (x1 - x0 * x0)
    n2 = golden["rosenbrock_f64"]["n2"]
    if abs(n2["f"] - f) > 1e-12 or abs(n2["grad"][0] - g0) > 1e-10 or abs(n2["grad"][1] - g1) > 1e-10:
        print("MISMATCH rosenbrock", f, g0, g1)
        bad += 1
    print("golden file consistent" if bad == 0 else f"{bad} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
