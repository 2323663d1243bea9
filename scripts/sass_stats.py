#!/usr/bin/env python
"""Static SASS statistics of the kernels (no GPU needed): registers per kernel and the instruction
mix of the innermost scene-record loop (the hot loop: from the first backward-branch target to the
branch).  Usage: python scripts/sass_stats.py [kernel-name-regex]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "svbrdf_estimation_b200", "csrc", "kernels.cu")
OUT = os.path.join(ROOT, "build", "kernels.cubin")


def rf_cycles(body):
    """Cost model measured on B200 (scripts/microbench3.cu): the register file delivers one 32-bit
    register per bank (even/odd) per cycle and SMSP, so an instruction occupies
    max(pipe cycles, distinct even source registers, distinct odd source registers) cycles; operands
    flagged .reuse, uniform registers and immediates are free; a packed F32x2 operand is an even+odd pair."""
    total = 0
    for _, op, args in body:
        base = op.split(".")[0]
        ops = [x.strip() for x in args.split(",")]
        srcs = ops[2:] if base in ("FSETP", "ISETP") else (ops if base == "BRA" else ops[1:])
        regs = set()
        for s in srcs:
            m = re.search(r"\bR(\d+)", s)
            if not m or ".reuse" in s:
                continue
            r = int(m.group(1))
            regs.add(r)
            if "F32x2" in s or ".64" in s:
                regs.add(r + 1)
        even = len([r for r in regs if r % 2 == 0])
        odd = len(regs) - even
        pipe = 2 if base in ("FFMA2", "FMUL2", "FADD2") else 1
        total += max(pipe, even, odd)
    return total


# the kernel RenderingLoss fwd+bwd launches on every power-of-two map size: packed lanes, BWD, grey lights, 12-channel
# layout, large parameter block, default accuracy, per-warp row table
DEFAULT_KERNEL = r"loss_kernel_packedINS_2F2ELb1ELb0ELb1ELi0ELi900ELb0ELb1EEE"


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    as_json = "--json" in sys.argv
    pat = re.compile(args[0] if args else DEFAULT_KERNEL)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    extra = os.environ.get("SVB_EXTRA_FLAGS", "").split()        # e.g. SVB_EXTRA_FLAGS="-DSVB_F2_MINB=2" for experiments
    p = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"] + extra +
                       ["-Xptxas", "-v", "-cubin", "-o", OUT, SRC], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode:
        print(p.stdout)
        return 1
    regs = {}
    cur = None
    for line in p.stdout.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), "spill" if "spill" in line else "")
        if "bytes spill" in line and cur and not line.strip().startswith("0 bytes stack frame, 0 bytes spill stores, 0 bytes spill loads"):
            print("SPILL in", cur, line.strip())
    sass = subprocess.run(["cuobjdump", "-sass", OUT], stdout=subprocess.PIPE, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if not pat.search(name):
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);", f)
        addr = {int(a, 16): i for i, (a, _, _) in enumerate(ins)}
        loops = []
        for i, (a, op, args) in enumerate(ins):
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", args)
                if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] < i:
                    loops.append((addr[int(m.group(1), 16)], i))
        if as_json:
            import json
            s0, e0 = [l for l in loops if any(op.startswith('MUFU') for _, op, _ in ins[l[0]:l[1] + 1])][0]   # first loop with MUFU work: the shared-roughness record loop
            body = ins[s0:e0 + 1]
            mix = collections.Counter(op.split(".")[0] for _, op, _ in body)
            packed_fma, packed_other = mix.get("FFMA2", 0), mix.get("FMUL2", 0) + mix.get("FADD2", 0)
            flop_pair = 4 * packed_fma + 2 * packed_other + 2 * mix.get("FFMA", 0) + mix.get("FMUL", 0) + mix.get("FADD", 0)
            print(json.dumps({"kernel": name, "registers": regs.get(name, (0, ""))[0], "instructions_total": len(ins),
                              "record_loop_instructions": len(body), "record_loop_mix": dict(mix.most_common()),
                              "executed_flop_per_eval": flop_pair / 2.0, "mufu_per_eval": mix.get("MUFU", 0) / 2.0,
                              "register_file_cycles_model": rf_cycles(body)}))
            continue
        print("== %s: %d registers, %d instructions total" % (name[:110], regs.get(name, (0, ""))[0], len(ins)))
        for (s, e) in loops:
            body = ins[s:e + 1]
            mix = collections.Counter(op.split(".")[0] for _, op, _ in body)
            fma_pipe = sum(v for k, v in mix.items() if k in ("FFMA", "FMUL", "FADD"))
            fma2 = sum(v for k, v in mix.items() if k in ("FFMA2", "FMUL2", "FADD2"))
            alu = sum(v for k, v in mix.items() if k in ("FMNMX", "FSEL", "FSETP", "LOP3", "IADD3", "MOV", "SEL", "ISETP", "PRMT", "SHF", "IMAD"))
            print("   loop [%d..%d] %d instr, ~%d register-file cycles: packed-FP %d, scalar-FP %d, MUFU %d, ALU-ish %d | %s" % (
                s, e, len(body), rf_cycles(body), fma2, fma_pipe, mix.get("MUFU", 0), alu,
                ", ".join("%s %d" % kv for kv in mix.most_common())))
    return 0


if __name__ == "__main__":
    sys.exit(main())
