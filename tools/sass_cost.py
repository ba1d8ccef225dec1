"""FP64-pipe cost model of a kernel's SASS (B200): per FP64 instruction max(2, number of distinct *vector*
register sources not served by the operand-reuse cache) cycles per warp per SMSP.  Measured rule
(tools/micro/dfma_operands.cu): DFMA with three distinct vector-register sources issues at 2/3 rate; uniform
registers, constant-bank operands, immediates and `.reuse`d operands are free.

    python tools/sass_cost.py <lib.so> <mangled-name-substring> [--loop-trip 8]
"""
import re, subprocess, sys

def kernel_sass(lib, pat):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    lines, on = [], False
    for l in out.splitlines():
        if "Function :" in l:
            on = pat in l
        elif on and re.search(r"/\*[0-9a-f]{4}\*/", l):
            lines.append(l)
    return lines

def analyse(lines, trip):
    addr = lambda l: int(re.search(r"/\*([0-9a-f]{4})\*/", l).group(1), 16)
    loops = []
    for l in lines:
        if "BRA" in l:
            m = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
            if m and int(m.group(1), 16) < addr(l):
                loops.append((int(m.group(1), 16), addr(l)))
    loop = max(loops, key=lambda x: x[1] - x[0]) if loops else (1 << 30, -1)
    res = {}
    for name, sel in (("loop", lambda a: loop[0] <= a <= loop[1]), ("outside", lambda a: not (loop[0] <= a <= loop[1]))):
        prev = {}
        n = cyc_noreuse = cyc_reuse = three = 0
        ops = {"DFMA": 0, "DADD": 0, "DMUL": 0}
        total = 0
        for l in lines:
            if not sel(addr(l)):
                continue
            total += 1
            m = re.search(r"\b(DFMA|DMUL|DADD)\s+(\S+),\s*(.*?);", l)
            if not m:
                continue
            ops[m.group(1)] += 1
            srcs = [o.strip() for o in m.group(3).split(",")]
            regs, new, cur = set(), 0, {}
            for slot, o in enumerate(srcs):
                r = re.match(r"[-|]*\s*(R\d+)(\.reuse)?", o)
                if not r or r.group(1) == "RZ":
                    continue
                regs.add(r.group(1))
                if prev.get(slot) != r.group(1):
                    new += 1
                if r.group(2):
                    cur[slot] = r.group(1)
            prev = cur
            n += 1
            cyc_noreuse += max(2, len(regs))
            cyc_reuse += max(2, min(new, len(regs)))
            three += len(regs) >= 3
        res[name] = dict(instrs=total, fp64=n, **ops, three_src=three, cycles_no_reuse=cyc_noreuse, cycles_with_reuse=cyc_reuse)
    tot_nr = res["loop"]["cycles_no_reuse"] * trip + res["outside"]["cycles_no_reuse"]
    tot_r = res["loop"]["cycles_with_reuse"] * trip + res["outside"]["cycles_with_reuse"]
    res["per_warp_cycles"] = dict(no_reuse=tot_nr, with_reuse=tot_r, fp64_instrs=res["loop"]["fp64"] * trip + res["outside"]["fp64"], all_instrs=res["loop"]["instrs"] * trip + res["outside"]["instrs"])
    return res

if __name__ == "__main__":
    lib, pat = sys.argv[1], sys.argv[2]
    trip = int(sys.argv[sys.argv.index("--loop-trip") + 1]) if "--loop-trip" in sys.argv else 8
    r = analyse(kernel_sass(lib, pat), trip)
    for k, v in r.items():
        print(k, v)
    # time at 128^3: warps per SMSP = 2097152/32/148/4
    w = 2097152 / 32 / 148 / 4
    for k in ("no_reuse", "with_reuse"):
        print(f"predicted FP64-pipe time at 128^3, 1.89 GHz ({k}): {r['per_warp_cycles'][k] * w / 1.89e9 * 1e3:.3f} ms")
