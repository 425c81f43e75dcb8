"""Static size of the sample loops of a kernel, from SASS (no GPU needed): every backward branch that encloses a 16-byte
gather (LDG.E.128) or a reduction (RED) is reported with its instruction count and opcode mix.

    python profiles/sass_loop_count.py vox-e_b200/csrc/build/voxe_render.o render_fwd_kernelILi0ELi3ELi96
"""
import re, subprocess, sys
def funcs(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur = None; d = {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m: cur = m.group(1); d[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur: d[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return d
def loops(ins):
    res = []
    for addr, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= addr:
                body = [(a, t) for a, t in ins if tgt <= a <= addr]
                res.append((tgt, addr, body))
    return res
obj, pat = sys.argv[1], sys.argv[2]
for name, ins in funcs(obj).items():
    if pat not in name: continue
    print(name[:90], "total", len(ins))
    for tgt, addr, body in loops(ins):
        n128 = sum("LDG.E.128" in t for _, t in body); red = sum(t.startswith("RED") or "REDG" in t or " RED" in t for _, t in body)
        if n128 or red:
            ops = {}
            for _, t in body:
                op = re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0]; ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda kv: -kv[1])[:12]
            print(f"  loop {tgt:#x}..{addr:#x}: {len(body)} instrs, LDG.128={n128}, RED={red}  {top}")
