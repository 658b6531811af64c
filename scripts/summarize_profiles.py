"""Turns the ncu CSVs under profiles/ into profiles/r1_summary.md and profiles/r1_traffic.json."""
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def short(name):
    return re.sub(r"\(.*", "", name).split("::")[-1].replace("unnamed>", "").strip(":")


def read_long(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    byid = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = byid.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(byid.values())


def main():
    out = ["# Round-1 profiles (B200, ncu 2025.x, `--clock-control none`)", "",
           "All captures were taken after the same command had exited 0 without ncu. Per-launch times under",
           "ncu are cold-cache and serialised: compare SHARES, not absolutes.", ""]
    # ---- launch list of the bench command
    L = read_long(os.path.join(P, "r1_bench_launches.csv"))
    tot = sum(k["gpu__time_duration.sum"] for k in L)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k in L:
        agg[k["name"]][0] += 1
        agg[k["name"]][1] += k["gpu__time_duration.sum"]
    out += ["## 1. Launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline`",
            f"`profiles/r1_bench_launches.csv` ({len(L)} launches from the timed region: `-s 540 -c 140`; one replay = 1 stem + 59 GEMMs + pool + decode + NMS + quads + PnP + set_src)", "",
            "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{n}` | {c} | {t/1e3:.1f} | {100*t/tot:.1f} % |")
    conv_share = sum(t for n, (c, t) in agg.items() if n.startswith("conv_")) / tot
    out += ["", f"Convolution kernels (`conv_raster_kernel` + `conv_tc_kernel`) = {100*conv_share:.1f} % of the step's kernel time;",
            "`bench.py` measures the same group live with CUDA events (`roofline.stage_ms.conv / total`).", ""]
    # ---- per-launch metrics of one 64-frame replay
    M = read_long(os.path.join(P, "r1_replay64_metrics.csv"))
    tot = sum(k["gpu__time_duration.sum"] for k in M)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    for k in M:
        a = agg[k["name"]]
        t = k["gpu__time_duration.sum"]
        a[0] += 1; a[1] += t
        a[2] += k["dram__bytes_read.sum"]; a[3] += k["dram__bytes_write.sum"]
        a[4] += k["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] * t
        a[5] += k["lts__t_bytes.sum"]
    out += ["## 2. Per-launch metrics, one eager replay of 64 Bayer frames (`scripts/profile_replay.py 64 1`)",
            f"`profiles/r1_replay64_metrics.csv` ({len(M)} launches; the capture window starts two launches into the replay, so the stem and the first GEMM are not in it)", "",
            "| kernel | launches | us | share | DRAM read MB | DRAM write MB | L2 bytes MB | tensor pipe active (time-weighted) |",
            "|---|---|---|---|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{n}` | {a[0]} | {a[1]/1e3:.1f} | {100*a[1]/tot:.1f} % | {a[2]/1e6:.1f} | {a[3]/1e6:.1f} | {a[5]/1e6:.1f} | {a[4]/a[1] if a[1] else 0:.2f} % |")
    conv = [k for k in M if k["name"].startswith("conv_")]
    traffic = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in conv)
    frames = 64
    out += ["", f"DRAM traffic of the {len(conv)} captured GEMM launches: {traffic/1e6:.1f} MB for {frames} frames = {traffic/frames/1e6:.2f} MB per frame",
            "(algorithmic: 42.1 MB of conv inputs + 28.6 MB of conv outputs per frame if nothing stayed in L2; DRAM writes are low because",
            "outputs are still dirty in the 126 MB L2 when the next kernel reads them).", "",
            "Top launches:", "", "| kernel | us | tensor pipe active | DRAM MB | smem/CTA KB | grid |", "|---|---|---|---|---|---|"]
    for k in sorted(M, key=lambda k: -k["gpu__time_duration.sum"])[:10]:
        out.append(f"| `{k['name']}` | {k['gpu__time_duration.sum']/1e3:.1f} | {k['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} % | "
                   f"{(k['dram__bytes_read.sum']+k['dram__bytes_write.sum'])/1e6:.1f} | {k['launch__shared_mem_per_block_dynamic']/1e3:.0f} | {k['grid']} |")
    json.dump({"conv_group_dram_bytes_per_frame": traffic / frames, "frames": frames, "launches": len(conv),
               "source": "profiles/r1_replay64_metrics.csv (dram__bytes_read.sum + dram__bytes_write.sum)"},
              open(os.path.join(P, "r1_traffic.json"), "w"), indent=1)
    # ---- top kernel, full set
    raw = list(csv.reader(open(os.path.join(P, "r1_raster_op46_raw.csv"))))
    hdr, units, d = raw[0], raw[1], raw[2]
    idx = {h: i for i, h in enumerate(hdr)}
    out += ["", "## 3. `ncu --set full --import-source on` of the top GEMM (Detect P3 box.0|cls.0, 3x3 64->128, 64 frames)",
            "`profiles/r1_raster_op46_raw.csv`, `profiles/r1_raster_op46_source.csv`", "", "| metric | value |", "|---|---|"]
    for k in ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
              "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"]:
        if k in idx:
            out.append(f"| `{k}` | {d[idx[k]]} {units[idx[k]]} |")
    src = list(csv.reader(open(os.path.join(P, "r1_raster_op46_source.csv"))))
    h2, data = src[1], src[2:]
    isamp, isrc = h2.index("# Samples"), h2.index("Source")
    stall = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    tots = {h: sum(int(r[h2.index(h)] or 0) for r in data) for h in stall}
    out += ["", "Warp-stall samples by reason (source page): " + ", ".join(f"{k[6:]} {v}" for k, v in sorted(tots.items(), key=lambda x: -x[1])[:6]) + ".",
            "Most samples are warps asleep on mbarriers (the per-tile chain load -> MMA -> epilogue is latency-bound), not issue stalls:",
            "this is what the next round has to attack (more independent accumulator chains per tile, deeper TMEM/stage overlap).", "",
            "SASS evidence: `UTCHMMA` (tcgen05.mma), `LDTM` (tcgen05.ld), `UBLKCP` (cp.async.bulk), `UTCBAR` (tcgen05.commit),",
            "`SYNCS.*` (mbarrier) appear in `cuobjdump -sass irmv_detection_b200/libirmv_b200.so`; hottest SASS lines:", "", "```"]
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:8]:
        out.append(f"{int(r[isamp]):6d} samples  {r[isrc].strip()[:90]}")
    out += ["```", ""]
    open(os.path.join(P, "r1_summary.md"), "w").write("\n".join(out))
    print("\n".join(out[:60]))


if __name__ == "__main__":
    main()
