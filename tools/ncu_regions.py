"""Stall samples of a kernel's prologue / main loop / epilogue from an ncu report (source page), plus headline metrics.

    python tools/ncu_regions.py gpurun_out/prof.ncu-rep
"""
import collections, csv, io, re, subprocess, sys


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:120])
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    src = [r[ix["Source"]] for r in data]
    addr = [int(r[ix["Address"]], 16) for r in data]
    loop = None
    for i, s in enumerate(src):
        m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            cand = [j for j, a in enumerate(addr) if (a - addr[0]) == tgt or a == tgt]
            if cand and cand[0] < i and (loop is None or i - cand[0] > loop[1] - loop[0]):
                loop = (cand[0], i)
    total = sum(int(r[ix["# Samples"]] or 0) for r in data)

    def region(a, b, name):
        t, n = collections.Counter(), 0
        for r in data[a : b + 1]:
            n += int(r[ix["# Samples"]] or 0)
            for s in stalls:
                if r[ix[s]]:
                    t[s[6:]] += int(r[ix[s]])
        fp = sum(1 for s in src[a : b + 1] if re.search(r"\b(DFMA|DADD|DMUL)\b", s))
        print(f"{name:9s} sass={b - a + 1:4d} fp64={fp:4d} samples={n:6d} ({100 * n / total:4.1f} %)  " + ", ".join(f"{k} {100 * v / total:.1f}" for k, v in t.most_common(7)))

    region(0, loop[0] - 1, "prologue")
    region(loop[0], loop[1], "loop")
    region(loop[1] + 1, len(data) - 1, "epilogue")
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    h, r = raw[0], raw[2]
    for k in ("gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads"):
        if k in h:
            print(f"  {k} = {r[h.index(k)]} {raw[1][h.index(k)]}")


if __name__ == "__main__":
    main(sys.argv[1])
