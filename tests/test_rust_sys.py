"""The Rust `-sys` crate (rust/eon-kzg-sys, uncompiled here: no rustc in the image) declares exactly the functions
of include/eon_kzg.h: the committed file equals what tools/gen_rust_sys.py generates from the header, every declared
function is exported by libeon_kzg.so, and the safe crate (rust/p3-eon-gpu) only calls functions that exist."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sys_crate_is_in_step_with_the_header():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_sys_crate_matches_library_exports():
    from plonky3_eon_b200 import lib
    handle = lib.load()
    text = open(os.path.join(ROOT, "rust", "eon-kzg-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (eon_[a-z0-9_]+)\(", text))
    assert declared == set(lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(handle, name), name


def test_safe_crate_calls_only_declared_functions():
    text = open(os.path.join(ROOT, "rust", "eon-kzg-sys", "src", "lib.rs")).read()
    declared = dict(re.findall(r"pub fn (eon_[a-z0-9_]+)\(\s*(.*?)\s*\)", text, flags=re.S))
    used = set()
    for f in ("lib.rs", "dft.rs", "pcs.rs"):
        src = open(os.path.join(ROOT, "rust", "p3-eon-gpu", "src", f)).read()
        for name, args in re.findall(r"sys::(eon_[a-z0-9_]+)\(\s*(.*?)\)\s*(?:}|;|\n)", src, flags=re.S):
            used.add(name)
            assert name in declared, f"{f} calls undeclared {name}"
            # same number of arguments as the declaration (top-level commas)
            depth, n = 0, 1 if args.strip() else 0
            for ch in args:
                depth += ch in "([{"
                depth -= ch in ")]}"
                n += ch == "," and depth == 0
            if args.rstrip().endswith(","):
                n -= 1
            want = len([a for a in declared[name].split(",") if a.strip()])
            assert n == want, f"{f}: {name} called with {n} arguments, declared with {want}"
    assert {"eon_mctx_create", "eon_mctx_kzg_commit", "eon_mctx_kzg_open_batch", "eon_mctx_coset_lde_batch"} <= used
