"""Writes tests/golden/cli_flags.json: every flag of the reference's parse_args() (train.py:37-99) with its default
and type, obtained by executing that function's own source (train.py itself cannot be imported here: its module
level imports pytorch_fid_wrapper).  TEST INFRASTRUCTURE, run in the build container only:

    python oracle/make_cli_fixture.py
"""
import argparse
import ast
import json
import os
import sys

REF = "/root/reference/train.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src = open(REF).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "parse_args")
    ns = {"argparse": argparse}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    argv, sys.argv = sys.argv, ["train.py"]
    try:
        defaults = vars(ns["parse_args"]())
    finally:
        sys.argv = argv
    flags = {k: {"default": v, "type": type(v).__name__} for k, v in sorted(defaults.items())}
    out = os.path.join(ROOT, "tests", "golden", "cli_flags.json")
    json.dump({"source": "train.py:37-99 parse_args()", "flags": flags}, open(out, "w"), indent=1)
    print("wrote", out, len(flags), "flags")


if __name__ == "__main__":
    main()
