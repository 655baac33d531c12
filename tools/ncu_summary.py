"""Dev tool: turn ncu outputs under gpurun_out/ into the markdown summaries committed under profiles/."""
import collections, csv, subprocess, sys


def launches(csv_path, out_path, title):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit, i_grid, i_blk = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.defaultdict(lambda: [0, 0.0, set()])
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[i_unit], 1e-6)
        k = r[i_name].split("(")[0].replace("void ", "")[:64]
        agg[k][0] += 1
        agg[k][1] += v
        agg[k][2].add(f"{r[i_grid]}x{r[i_blk]}")
    tot = sum(v[1] for v in agg.values())
    with open(out_path, "w") as f:
        f.write(f"# {title}\n\nPer-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n\n| kernel | launches | total ms | share | grid x block |\n|---|---:|---:|---:|---|\n")
        for k, (n, t, g) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {t:.2f} | {100 * t / tot:.1f} % | {', '.join(sorted(g))[:60]} |\n")


def full(rep_path, out_path, title, top=25, kernel=""):
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    i_kn = hdr.index("Kernel Name")
    vals = [r for r in rows[2:] if kernel in r[i_kn]][-1]  # last launch of the wanted kernel
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    src = subprocess.run(["ncu", "-i", rep_path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows_all = list(csv.reader(src.splitlines()))
    # the source page has one section per profiled launch, each introduced by a "Kernel Name" row
    starts = [i for i, r in enumerate(srows_all) if r and r[0] == "Kernel Name"]
    pick = [i for i in starts if kernel in srows_all[i][1]][-1] if starts else 0
    end = min([i for i in starts if i > pick] + [len(srows_all)])
    srows = srows_all[pick:end]
    with open(out_path, "w") as f:
        f.write(f"# {title}\n\n| metric | unit | value |\n|---|---|---|\n")
        for h, u, v in zip(hdr, units, vals):
            if h in want:
                f.write(f"| {h} | {u} | {v} |\n")
        if len(srows) > 3:
            sh = srows[1]
            idx = {h: i for i, h in enumerate(sh)}
            data = [r for r in srows[2:] if len(r) == len(sh)]
            S = idx["# Samples"]
            tot = sum(int(r[S] or 0) for r in data) or 1
            f.write(f"\n## warp-stall sampling ({tot} samples)\n\n| reason | share |\n|---|---:|\n")
            st = {h: sum(int(r[idx[h]] or 0) for r in data) for h in sh if h.startswith("stall_") and "Not Issued" not in h}
            for h, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
                f.write(f"| {h} | {100 * v / tot:.1f} % |\n")
            f.write(f"\n## hottest SASS instructions\n\n| samples | executed | SASS |\n|---:|---:|---|\n")
            for r in sorted(data, key=lambda r: -int(r[S] or 0))[:top]:
                f.write(f"| {r[S]} | {r[idx['Instructions Executed']]} | `{r[1][:80]}` |\n")


def regions(rep_path, out_path, title, warps_per_pass=0):
    """Splits the SASS of the (single) profiled kernel at its BAR.SYNC instructions and lists, per region, the static
    and executed instruction counts, the share of the stall samples and the top stall reasons - the view that showed
    which phases of the persistent kernel are issue bound (executed instructions per element) and which wait."""
    src = subprocess.run(["ncu", "-i", rep_path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, data = rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    regs, cur = [], None
    for r in data:
        if len(r) != len(hdr):
            continue
        if cur is None:
            cur = {"samples": 0, "st": collections.Counter(), "instr": 0, "exec": 0, "wf": 0, "ideal": 0}
        cur["samples"] += int(r[col["# Samples"]] or 0)
        cur["instr"] += 1
        cur["exec"] += int(r[col["Instructions Executed"]] or 0)
        cur["wf"] += int(r[col["L1 Wavefronts Shared"]] or 0)
        cur["ideal"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
        for h in stalls:
            v = int(r[col[h]] or 0)
            if v:
                cur["st"][h[6:]] += v
        if "BAR.SYNC" in r[col["Source"]]:
            regs.append(cur)
            cur = None
    if cur:
        regs.append(cur)
    tot = sum(x["samples"] for x in regs) or 1
    with open(out_path, "w") as f:
        f.write(f"# {title}\n\nRegions = stretches of SASS between two BAR.SYNC instructions, in program order; regions with < 0.5 % of the samples omitted.\n\n")
        f.write("| region | SASS instr | executed (warp-level) | samples | share | shared wavefronts / ideal | top stall reasons |\n|---:|---:|---:|---:|---:|---|---|\n")
        for i, x in enumerate(regs):
            if x["samples"] < 0.005 * tot:
                continue
            top = ", ".join(f"{k} {v}" for k, v in x["st"].most_common(4))
            f.write(f"| {i} | {x['instr']} | {x['exec']} | {x['samples']} | {100 * x['samples'] / tot:.1f} % | {x['wf']} / {x['ideal']} | {top} |\n")


if __name__ == "__main__":
    kind = sys.argv[1]
    if kind == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif kind == "regions":
        regions(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], kernel=sys.argv[5] if len(sys.argv) > 5 else "")
