#!/usr/bin/env python
"""Compile the reference's .skel assets into dart_env_b200/assets/*.model.json.

/root/reference does not exist on the GPU box, so the four in-scope models are shipped as
compiled, engine-agnostic model descriptions (the parser's output, not the XML).  Run here:
    python tools/export_models.py
tests/test_skel_and_boundary.py::test_bundled_models_match_skel_files keeps them in sync."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dart_env_b200.skel import find_asset, parse_skel  # noqa: E402
from dart_env_b200.tasks import SPECS  # noqa: E402

EXTRA = ["cartpole.skel", "cartpole_swingup.skel", "inverted_double_pendulum.skel", "reacher2d.skel"]
for skel in [spec.skel for spec in SPECS.values()] + EXTRA:
    class spec:  # noqa: N801
        pass
    spec.skel = skel
    src = find_asset(spec.skel)
    m = parse_skel(src)  # file's own <time_step>; the env passes dt at load time
    out = os.path.join(ROOT, "dart_env_b200", "assets", spec.skel[:-5] + ".model.json")
    m.save_json(out)
    print(src, "->", out, os.path.getsize(out), "bytes")
