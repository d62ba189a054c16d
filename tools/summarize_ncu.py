"""Turns the ncu outputs in gpurun_out/ into the text summaries committed under profiles/:
  python tools/summarize_ncu.py <tag>          (expects gpurun_out/{launches,decode,train,march}_<tag>.*)
Writes profiles/launches_<tag>.md (kernel shares of one profiling pass) and profiles/<kernel>_<tag>.md
(key metrics of `ncu --set full` + warp-stall breakdown)."""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def short(n):
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("vnr::", "")[:48]


def launches():
    f = os.path.join(G, f"launches_{tag}.csv")
    if not os.path.exists(f):
        return
    lines = [l for l in open(f) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"]) + "  grid " + r["Grid Size"].replace(" ", "") + " block " + r["Block Size"].replace(" ", "")
        v = float(r["Metric Value"].replace(",", ""))
        a = tot.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(P, f"launches_{tag}.md"), "w") as o:
        o.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over tools/profile_render.py\n\n")
        o.write("100 training steps (batch 2^16) + 3 frames 1024^2 (host-enqueued rounds, VNR_RM_GRAPH=0) + 4 training steps (batch 2^18).\n")
        o.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
            o.write(f"| {k} | {c} | {t/1e3:.1f} | {t/c/1e3:.2f} | {100*t/total:.1f} % |\n")
        # one frame in detail
        idx = [i for i, r in enumerate(rows) if "march_round_kernel<(bool)1>" in r["Kernel Name"] or "march_round_kernel<1>" in r["Kernel Name"]]
        if len(idx) >= 2:
            s, e = idx[1], (idx[2] if len(idx) > 2 else len(rows))
            fr = rows[s:e]
            ft = sum(float(r["Metric Value"].replace(",", "")) for r in fr)
            agg = collections.OrderedDict()
            for r in fr:
                v = float(r["Metric Value"].replace(",", ""))
                a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
            o.write(f"\n## one frame (second of three): {len(fr)} launches, {ft/1e3:.1f} us serialised\n\n| kernel | launches | total us | largest us | share of frame |\n|---|---|---|---|---|\n")
            for k, (c, t, m) in sorted(agg.items(), key=lambda x: -x[1][1]):
                o.write(f"| {k} | {c} | {t/1e3:.1f} | {m/1e3:.1f} | {100*t/ft:.1f} % |\n")
    print("wrote launches")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source_regions(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next((r for r in rows if "Source" in r and "# Samples" in r), None)
    if not hdr:
        return None
    ia, isamp = hdr.index("Source"), hdr.index("# Samples")
    data = [r for r in rows if len(r) > max(ia, isamp) and r[isamp].isdigit()]
    n = len(data) // 2 if len(data) > 10 and data[0][ia] == data[len(data) // 2][ia] else len(data)   # the page is emitted twice
    data = data[:n]
    tot = sum(int(r[isamp]) for r in data) or 1
    segs, acc, start = [], 0, 0
    for i, r in enumerate(data):
        s = r[ia].strip(); parts = s.split()
        op = parts[0] if not s.startswith("@") else (parts[1] if len(parts) > 1 else "")
        acc += int(r[isamp])
        if op.startswith(("BAR", "UTCBAR", "SYNCS", "BRA", "EXIT", "RED", "ATOM", "WARPSYNC", "UTCHMMA")):
            if acc > 0.02 * tot:
                segs.append((start, i + 1, acc, s[:70]))
            acc = 0; start = i + 1
    top = sorted(data, key=lambda r: -int(r[isamp]))[:12]
    return tot, segs, [(int(r[isamp]), r[ia].strip()[:90]) for r in top]


def kernel_report(name, rep, regex):
    if not os.path.exists(rep):
        return
    hdr, units, rows = raw(rep)
    with open(os.path.join(P, f"{name}_{tag}.md"), "w") as o:
        o.write(f"# ncu --set full --clock-control none --import-source on: {name} ({tag})\n\nSource: gpurun_out/{os.path.basename(rep)} (scratch, not committed); one column per captured launch.\n\n")
        names = [short(r[hdr.index("Kernel Name")]) for r in rows]
        o.write("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(rows) + "\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                o.write(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows) + " |\n")
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        o.write("\n## warp stall reasons (warps stalled per issue-active cycle)\n\n| reason | " + " | ".join(names) + " |\n|---|" + "---|" * len(rows) + "\n")
        for i, h in sorted(stall, key=lambda x: -max(float(r[x[0]].replace(",", "") or 0) for r in rows)):
            vals = [r[i] for r in rows]
            if max(float(v.replace(",", "") or 0) for v in vals) < 0.05:
                continue
            o.write(f"| {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | " + " | ".join(vals) + " |\n")
        sr = source_regions(rep, regex)
        if sr:
            tot, segs, top = sr
            o.write(f"\n## PC-sampling by SASS region (first captured launch; {tot} samples; regions end at the named instruction)\n\n| SASS lines | samples | share | region ends at |\n|---|---|---|---|\n")
            for a, b, c, s in segs:
                o.write(f"| {a}-{b} | {c} | {100*c/tot:.1f} % | `{s}` |\n")
            o.write("\nhottest instructions:\n\n")
            for c, s in top:
                o.write(f"- {c} ({100*c/tot:.1f} %) `{s}`\n")
    print("wrote", name)


def decode_traffic(rep_name="decode", log_name="prof_decode.log", cfg="t19_1024x1024"):
    """profiles/decode_traffic_<config>.json: DRAM bytes of the first captured decode launch (round 0 of frame 0) next to its sample count
    (the per-round counters tools/profile_render.py printed in the same run) -- bench.py's roofline.traffic."""
    import json
    rep, log = os.path.join(G, f"{rep_name}_{tag}.ncu-rep"), os.path.join(G, log_name)
    if not (os.path.exists(rep) and os.path.exists(log)):
        return
    m = re.search(r"round_counts 0 \[(\d+)", open(log).read())
    if not m:
        return
    hdr, units, rows = raw(rep)
    r = rows[0]
    def val(k):
        i = hdr.index(k); v = float(r[i].replace(",", "")); u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
    dur = float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")); du = units[hdr.index("gpu__time_duration.sum")]
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0)
    out = {"kernel": short(r[hdr.index("Kernel Name")]), "samples": int(m.group(1)), "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"), "duration_us": dur_us,
           "source": f"profiles/decode_{tag}.md (ncu --set full, first decode launch of frame 0 = wavefront round 0)"}
    out["config"] = cfg
    out["source"] = f"gpurun_out/{rep_name}_{tag}.ncu-rep (ncu --set full, first decode launch of frame 0 = wavefront round 0)"
    for k in ("lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"):
        if k in hdr:
            out[k] = r[hdr.index(k)]
    json.dump(out, open(os.path.join(P, f"decode_traffic_{cfg}.json"), "w"), indent=1)
    print(f"wrote decode_traffic_{cfg}.json", out)


launches()
decode_traffic()
decode_traffic("decode4k", "prof_decode4k.log", "t22_3840x2160")
kernel_report("decode", os.path.join(G, f"decode_{tag}.ncu-rep"), "decode_kernel")
kernel_report("train", os.path.join(G, f"train_{tag}.ncu-rep"), "train_step")
kernel_report("march", os.path.join(G, f"march_{tag}.ncu-rep"), "march_round")
