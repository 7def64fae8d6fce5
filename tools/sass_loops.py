"""Static view of a kernel's loops: every backward branch of one function of a built library with
the size of its body and the opcode mix inside (no GPU needed).

usage: python tools/sass_loops.py LIB.so FUNCTION_SUBSTRING [min_body [max_body [--dump]]]
e.g.   python tools/sass_loops.py densepoints_b200/_build/libdensepoints_cuda.so \
           dp_refine_group_kernelI10DpGroupCfgILi4ELi13E 40
"""
import collections
import re
import subprocess
import sys


def function_sass(lib, needle):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    ins, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = needle in line
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def opcode(text):
    text = re.sub(r"^@!?U?P\w+\s+", "", text)
    return text.split()[0]


def main():
    lib, needle = sys.argv[1], sys.argv[2]
    min_body = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    max_body = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 30
    dump = "--dump" in sys.argv
    ins = function_sass(lib, needle)
    if not ins:
        raise SystemExit("function not found")
    addr = [a for a, _ in ins]
    print(f"{needle}: {len(ins)} instructions")
    for k, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\w+,\s*)?(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr:
            continue
        body = ins[addr.index(tgt):k + 1]
        if len(body) < min_body or len(body) > max_body:
            continue
        mix = collections.Counter(opcode(x).split(".")[0] for _, x in body)
        top = " ".join(f"{o}:{n}" for o, n in mix.most_common(14))
        print(f"  loop 0x{tgt:04x}..0x{a:04x}  {len(body):4d} instr | {top}")
        if dump:
            for ba, bt in body:
                print(f"      {ba:04x}  {bt}")


if __name__ == "__main__":
    main()
