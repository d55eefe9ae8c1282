#!/usr/bin/env python
"""Time single stages in isolation (CUDA events on the context stream): python tools/time_stage.py [W H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vk_gltf_viewer_b200 import api
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
r = api.Renderer(W, H)
def t(fn, n=50, flush=False):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(n):
        if flush: r.flush_l2()
        r.event_record(0); fn(); r.event_record(1)
        tot += r.event_elapsed(0, 1)
    return tot / n * 1e3
print(f"{W}x{H}: clear {t(r.clear):.1f} us (L2-flushed {t(r.clear, flush=True):.1f}), hiz {t(r.hiz):.1f} us (L2-flushed {t(r.hiz, flush=True):.1f})")
