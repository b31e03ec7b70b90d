#!/usr/bin/env python
"""usage: tools/build_variant.py <name> [-DMACRO=VALUE ...]   -> build/variants/lib_<name>.so
A variant build of the library for kernel A/B runs on the GPU box (MUSE_B200_LIB=build/variants/lib_<name>.so)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "go-muse_b200"))
import muse_b200  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
out = os.path.join(ROOT, "build", "variants", "lib_%s.so" % name)
print(muse_b200.build(verbose=False, out=out, extra_flags=flags, tag=name))
