#!/usr/bin/env python3
"""Mnemonic counts per kernel from `cuobjdump -sass` of the shipped sm_100a objects (colate_b200/csrc/*.o).
usage: python tools/sass_table.py > table.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "colate_b200", "csrc")
KERNELS = ["k_ok", "k_replay", "k_emp", "k_sample", "k_compact", "k_join", "k_pileup_join", "k_em_cta", "k_em_split", "k_em", "k_bootstrap", "k_gen_tma",
           "k_gen", "k_jump", "k_decode_colate_in", "k_parse_mut"]
COLS = [("UBLKCP", r"\bUBLKCP"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"), ("STAS", r"\bSTAS"), ("UCGABAR", r"\bUCGABAR"), ("DFMA", r"\bDFMA"),
        ("DADD", r"\bDADD"), ("DMUL", r"\bDMUL"), ("MUFU.LG2", r"MUFU\.LG2"), ("I2F.F64", r"I2F\.F64"), ("REDS/ATOMS", r"\b(REDS|ATOMS)"),
        ("IDP.4A", r"IDP\.4A"), ("SHFL", r"\bSHFL"), ("LDS.128", r"LDS\.128"), ("LOP3", r"\bLOP3"), ("BAR.SYNC", r"BAR\.SYNC")]
print("| object | kernel | instructions | " + " | ".join(c for c, _ in COLS) + " |")
print("|---|---|---|" + "---|" * len(COLS))
for obj in ("kernels_sites.o", "kernels_em.o", "kernels_mt.o", "kernels_ingest.o"):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)[1:]
    for f in funcs:
        mangled = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", mangled], capture_output=True, text=True).stdout.strip() or mangled
        name = re.sub(r"\(.*", "", dem.replace("(bool)", "")).replace("colate::", "").replace("void ", "")
        base = re.sub(r"<.*", "", name)
        if base not in KERNELS:
            continue
        lines = [ln for ln in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        print("| %s | %s | %d | %s |" % (obj, name, len(lines), " | ".join(str(sum(1 for ln in lines if re.search(p, ln))) for _, p in COLS)))
