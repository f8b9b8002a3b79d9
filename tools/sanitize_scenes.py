#!/usr/bin/env python
"""Small scenes for compute-sanitizer (contact with friction, cloth over a scripted box, foam block under hinged boxes):
    compute-sanitizer --tool memcheck|racecheck|synccheck --error-exitcode 7 python tools/sanitize_scenes.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stark_b200 import scenes
for name, n, steps in (("tetdrop", 4, 6), ("cloth_shells", 8, 14), ("tetchain", 3, 6)):
    kw = {"ny": 4} if name == "tetchain" else {}
    sc = scenes.Scene(name, n=n, **kw)
    its = 0
    for _ in range(steps):
        s = sc.step(); its += int(s["newton_iterations"])
    print(name, "its", its, flush=True)
