"""Per-function hot spots of one kernel: ncu SASS page (executed instructions, thread instructions, stall samples per
address) joined with `nvdisasm -gi` inline chains, aggregated INCLUSIVELY by the device function each instruction was
inlined from.

usage: python scripts/ncu_hotspots.py <report.ncu-rep> <object.o> <mangled-substring>|<demangled-substring> [source.cu ...]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def function_spans(path):
    """[(first_line, last_line, name)] of the __device__/__global__ functions of a source file (brace matching)."""
    text = open(path).read().split("\n")
    spans = []
    i = 0
    sig = re.compile(r"(__device__|__global__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    while i < len(text):
        m = sig.search(text[i])
        if m and not text[i].lstrip().startswith("//"):
            name = m.group(2)
            depth = 0
            started = False
            j = i
            while j < len(text):
                for ch in text[j]:
                    if ch == "{":
                        depth += 1
                        started = True
                    elif ch == "}":
                        depth -= 1
                if started and depth == 0:
                    break
                j += 1
            spans.append((i + 1, j + 1, name))
            i = j + 1
        else:
            i += 1
    return spans


def main():
    rep, obj, kname = sys.argv[1:4]
    kname, _, kdemangled = kname.partition("|")
    kdemangled = kdemangled or kname
    sources = sys.argv[4:]
    spans = {os.path.abspath(s): function_spans(s) for s in sources}

    def func_of(path, line):
        for a, b, n in spans.get(os.path.abspath(path), []):
            if a <= line <= b:
                return n
        return None

    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    chains = {}   # offset -> [function names, innermost first]
    opcodes = {}
    in_kernel = False
    chain = []
    pending = []
    for ln in dis:
        if ln.startswith("\t.section\t.text."):
            in_kernel = kname in ln
            pending = []
            continue
        if not in_kernel:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            if pending:
                chain = []
                for path, line in pending:
                    f = func_of(path, line)
                    if f and (not chain or chain[-1] != f):
                        chain.append(f)
                pending = []
            off = int(m.group(1), 16)
            chains[off] = chain
            ins = m.group(2).strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            opcodes[off] = ins.split()[0].split(".")[0]

    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and kdemangled in r[1]:
            start = i
            break
    if start is None:
        raise SystemExit("kernel not found in report")
    hdr = rows[start + 1]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    incl_inst = defaultdict(int)
    incl_thr = defaultdict(int)
    incl_smp = defaultdict(int)
    self_inst = defaultdict(int)
    op_inst = defaultdict(int)
    op_smp = defaultdict(int)
    tot_inst = tot_thr = tot_smp = 0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stalls = defaultdict(int)
    nsass = defaultdict(int)
    for r in rows[start + 2:]:
        if not r or r[0] == "Kernel Name":
            break
        addr = int(r[col["Address"]], 16)
        if base is None:
            base = addr
        off = addr - base
        inst = int(r[col["Instructions Executed"]] or 0)
        thr = int(r[col["Thread Instructions Executed"]] or 0)
        smp = int(r[col["# Samples"]] or 0)
        tot_inst += inst
        tot_thr += thr
        tot_smp += smp
        ch = chains.get(off, [])
        for f in set(ch):
            incl_inst[f] += inst
            incl_thr[f] += thr
            incl_smp[f] += smp
            nsass[f] += 1
        self_inst[ch[0] if ch else "?"] += inst
        op_inst[opcodes.get(off, "?")] += inst
        op_smp[opcodes.get(off, "?")] += smp
        for h in stall_cols:
            stalls[h] += int(r[col[h]] or 0)
    print(f"kernel {kname}: {tot_inst} warp-instructions, {tot_thr / max(tot_inst, 1):.2f} active threads/instruction, {tot_smp} samples, {len(chains)} SASS instructions")
    print("\ninclusive by device function (an instruction counts for every function on its inline chain)")
    print(f"{'function':28s} {'inst%':>7s} {'samples%':>9s} {'thr/inst':>9s} {'SASS':>6s} {'self inst%':>10s}")
    for f, v in sorted(incl_inst.items(), key=lambda kv: -kv[1]):
        print(f"{f:28s} {100 * v / tot_inst:7.1f} {100 * incl_smp[f] / max(tot_smp, 1):9.1f} {incl_thr[f] / max(v, 1):9.1f} {nsass[f]:6d} {100 * self_inst.get(f, 0) / tot_inst:10.1f}")
    print("\nby opcode")
    for o, v in sorted(op_inst.items(), key=lambda kv: -kv[1])[:24]:
        print(f"{o:12s} inst {100 * v / tot_inst:5.1f}%  samples {100 * op_smp[o] / max(tot_smp, 1):5.1f}%")
    print("\nstall samples")
    for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"{h:28s} {100 * v / max(tot_smp, 1):5.1f}%")


if __name__ == "__main__":
    main()
