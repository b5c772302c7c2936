#!/usr/bin/env python
"""Experiment builds of libiris under gpurun_scratch/<name>/ (they travel to the GPU box with the
snapshot): name=FLAG[,FLAG...] per argument, e.g.  trace=-DIRIS_TRACE  fr9=-DIRIS_FIX_FR=9"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from challenge_b200 import build as b

for spec in sys.argv[1:]:
    name, _, flags = spec.partition('=')
    d = os.path.join(ROOT, 'gpurun_scratch', name)
    print(b.build(extra_flags=[f for f in flags.split(',') if f], lib=os.path.join(d, 'libiris.so'),
                  obj_dir=os.path.join(d, 'obj')))
