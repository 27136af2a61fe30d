#!/usr/bin/env python
"""Diff two directories of ShaderNN per-layer `.dump` files (one written by `snnb_model_dump_outputs`, the other e.g. by a
ShaderNN / ncnn run of the same model elsewhere) with the reference's comparator.
Usage: python tools/compare_dumps.py DIR_A DIR_B [--eps 0.01]        exit status 1 if any layer mismatches."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shadernn_b200 import dumpio  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir_a")
    ap.add_argument("dir_b")
    ap.add_argument("--eps", type=float, default=0.01, help="abs-AND-rel tolerance of the reference's comparator (testutil.cpp:351-361)")
    args = ap.parse_args()
    rows = dumpio.compare_dirs(args.dir_a, args.dir_b, args.eps)
    bad = 0
    for name, shape, mism, mx, status in rows:
        print("%-64s %-16s %10s %12s  %s" % (name, "" if shape is None else "x".join(map(str, shape)), "" if mism is None else mism,
                                           "" if mx is None else "%.3e" % mx, status))
        bad += status != "ok"
    print("%d file(s), %d not ok" % (len(rows), bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
