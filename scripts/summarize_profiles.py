"""Turns the ncu CSVs under profiles/ into profiles/r1_summary.md and profiles/r1_traffic.json.
The CSVs come from scripts/collect_profiles.sh (run on the GPU box, copied from gpurun_out/)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from analyze_launches import name_ops  # noqa: E402


def short(name):
    return re.sub(r"\(.*", "", name).split("::")[-1].split("<")[0].replace("unnamed>", "").strip(":")


def read_long(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    byid = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = byid.setdefault(r["ID"], {"name": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return list(byid.values())


def full_capture(out, tag, title, keys):
    raw = list(csv.reader(open(os.path.join(P, f"{tag}_raw.csv"))))
    hdr, units, d = raw[0], raw[1], raw[2]
    idx = {h: i for i, h in enumerate(hdr)}
    out += ["", title, f"`profiles/{tag}_raw.csv`, `profiles/{tag}_source.csv`", "", "| metric | value |", "|---|---|"]
    for k in keys:
        if k in idx:
            out.append(f"| `{k}` | {d[idx[k]]} {units[idx[k]]} |")
    src = list(csv.reader(open(os.path.join(P, f"{tag}_source.csv"))))
    h2, data = src[1], src[2:]
    isamp, isrc, iex = h2.index("# Samples"), h2.index("Source"), h2.index("Instructions Executed")
    stall = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    tots = {h: sum(int(r[h2.index(h)] or 0) for r in data) for h in stall}
    ops = collections.Counter()
    for r in data:
        t = r[isrc].split()
        op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "?")).split(".")[0]
        ops[op] += int(r[iex] or 0)
    tot_i = sum(ops.values())
    out += ["", "Warp-stall samples by reason (source page): " + ", ".join(f"{k[6:]} {v}" for k, v in sorted(tots.items(), key=lambda x: -x[1])[:6]) + ".",
            "Executed instructions by opcode: " + ", ".join(f"{k} {100 * v / tot_i:.1f} %" for k, v in ops.most_common(8)) + ".", "",
            "Hottest SASS lines:", "", "```"]
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:8]:
        out.append(f"{int(r[isamp]):6d} samples  {r[isrc].strip()[:90]}")
    out += ["```"]
    return {h: d[idx[h]] for h in keys if h in idx}


def main():
    out = ["# Round-1 profiles (B200, ncu 2025.x, `--clock-control none`)", "",
           "Produced by `scripts/collect_profiles.sh` on a B200 (every capture after the same command had exited 0 without ncu)",
           "and summarised by `scripts/summarize_profiles.py`.  Per-launch times under ncu are cold-cache and serialised:",
           "compare SHARES, not absolutes.", ""]
    # ---- launch list of the bench command
    L = read_long(os.path.join(P, "r1_bench_launches.csv"))
    starts = [i for i, k in enumerate(L) if k["name"] == "set_src_kernel"]
    step = L[starts[0]:starts[2]]                 # two 128-frame replays = one 256-frame step
    tot = sum(k["gpu__time_duration.sum"] for k in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k in step:
        agg[k["name"]][0] += 1
        agg[k["name"]][1] += k["gpu__time_duration.sum"]
    out += ["## 1. Launch list of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline`",
            f"`profiles/r1_bench_launches.csv` ({len(L)} launches of the engine's kernels, `-s 374 -c 420`; the table is one step = two",
            f"128-frame replays = {len(step)} launches: set_src + stem + {sum(1 for k in step if k['name'].startswith('conv_')) // 2} GEMM launches + pool + decode + NMS + quads + PnP per replay;",
            "eleven 1x1 convs run as fused tails of their producers, conv0 inside the stem)", "",
            "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{n}` | {c} | {t/1e3:.1f} | {100*t/tot:.1f} % |")
    conv_share = sum(t for n, (c, t) in agg.items() if n.startswith("conv_")) / tot
    out += ["", f"Convolution kernels (`conv_raster_kernel` + `conv_tc_kernel`) = {100*conv_share:.1f} % of the step's kernel time;",
            "`bench.py` measures the same group live with CUDA events (`roofline.stage_ms.conv / total`).", ""]
    # ---- per-launch metrics of one 128-frame replay
    frames = 128
    M = read_long(os.path.join(P, "r1_replay128_metrics.csv"))
    tot = sum(k["gpu__time_duration.sum"] for k in M)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    for k in M:
        a = agg[k["name"]]
        t = k["gpu__time_duration.sum"]
        a[0] += 1; a[1] += t
        a[2] += k["dram__bytes_read.sum"]; a[3] += k["dram__bytes_write.sum"]
        a[4] += k["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] * t
        a[5] += k["lts__t_bytes.sum"]
    out += [f"## 2. Per-launch counters, one eager replay of {frames} Bayer frames (`scripts/ncu_replay_metrics.sh 128`)",
            f"`profiles/r1_replay128_metrics.csv` ({len(M)} launches = the whole replay)", "",
            "| kernel | launches | us | share | DRAM read MB | DRAM write MB | L2 bytes MB | tensor pipe active (time-weighted) |",
            "|---|---|---|---|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{n}` | {a[0]} | {a[1]/1e3:.1f} | {100*a[1]/tot:.1f} % | {a[2]/1e6:.1f} | {a[3]/1e6:.1f} | {a[5]/1e6:.1f} | {a[4]/a[1] if a[1] else 0:.2f} % |")
    conv = [k for k in M if k["name"].startswith("conv_")]
    traffic = sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in conv)
    conv_us = sum(k["gpu__time_duration.sum"] for k in conv) / 1e3
    out += ["", f"DRAM traffic of the {len(conv)} GEMM launches: {traffic/1e6:.1f} MB for {frames} frames = {traffic/frames/1e6:.2f} MB per frame in {conv_us:.0f} us",
            f"= {traffic / conv_us / 1e6:.2f} TB/s averaged over the conv stage (algorithmic: 42.1 MB of conv inputs + 28.6 MB of conv outputs per frame when",
            "nothing stays in the 126 MB L2; at 128 frames per replay the 160x160 and 80x80 tensors do not fit, so m1-m4 run at the HBM roofline).", ""]
    convs = [k for k in M if k["name"].startswith("conv_") or k["name"].startswith("sppf")]
    LY = name_ops(json.load(open(os.path.join(P, "r1_ops.json"))))
    out += ["Per layer (network order; `hw` = output side, tensor % = `sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed`):", "",
            "| layer | kernel | hw | cin | cout | k | s | us | TFLOP/s | tensor % | DRAM MB | DRAM TB/s | L2 MB | smem KB |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    if len(convs) == len(LY):
        for k, d in zip(convs, LY):
            name, hw, cin, cout, kk, s, flf = d
            t = k["gpu__time_duration.sum"] / 1e3
            if name == "POOL":
                out.append(f"| SPPF pool | `{k['name']}` | 20 | 128 | 384 | 5 | 1 | {t:.1f} | | | | | | |")
                continue
            fl = flf * frames
            dm = (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) / 1e6
            out.append(f"| {name} | `{k['name'].replace('_kernel', '')}` | {hw} | {cin} | {cout} | {kk} | {s} | {t:.1f} | {fl/t/1e6:.0f} | "
                       f"{k['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} | {dm:.0f} | {dm/t:.2f} | {k['lts__t_bytes.sum']/1e6:.0f} | "
                       f"{k['launch__shared_mem_per_block_dynamic']/1e3:.0f} |")
    json.dump({"conv_group_dram_bytes_per_frame": traffic / frames, "frames": frames, "launches": len(conv),
               "source": "profiles/r1_replay128_metrics.csv (dram__bytes_read.sum + dram__bytes_write.sum of the GEMM launches)"},
              open(os.path.join(P, "r1_traffic.json"), "w"), indent=1)
    # ---- full captures
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    h0 = full_capture(out, "r1_raster_h0", "## 3. `ncu --set full --import-source on` of the top GEMM (Detect P3 box.0|cls.0: 3x3, 64 -> 128, 128 frames)", keys)
    out += ["", "The MMA issue loop now runs at the tensor pipe's own rate (scripts/mma_probe.cu: max(N/2, 32 + N/4) cycles per M128 x N x K16 MMA,",
            "the 32 + N/4 term being the shared-memory operand fetch); what is left is the per-tile hand-off (a two-stage activation ring next to",
            "147 KB of resident weights) and the epilogue's TMEM round trip."]
    full_capture(out, "r1_stem", "## 4. `ncu --set full --import-source on` of `stem_kernel` (demosaic + rot180 + resize + conv0, 128 Bayer frames)", keys)
    out += ["", "The stem is instruction-issue bound (`smsp__issue_active` above), not HBM bound: 1.31 MB in + 3.28 MB out per frame would take",
            "0.7 us at the measured HBM peak.", "",
            "## 5. SASS evidence", "",
            "`cuobjdump -sass irmv_detection_b200/libirmv_b200.so` contains `UTCHMMA` (tcgen05.mma), `LDTM` (tcgen05.ld), `UBLKCP`",
            "(cp.async.bulk), `UTCBAR` (tcgen05.commit), `SYNCS.*` (mbarrier), `ACQBULK`/`griddepcontrol` (programmatic dependent launch)",
            "and `HMMA.16816` (the stem's conv0).", ""]
    if os.path.exists(os.path.join(P, "r1_armors_raw.csv")):
        full_capture(out, "r1_armors", "## 6. `ncu --set full --import-source on` of `extract_armors_kernel` (64 frames x 10 detections, "
                     "seeded light-bar scenes, `scripts/bench_armors.py`)", keys)
        out += ["", "One CTA per detection; the border walks are serial per component (one lane), so the kernel is latency bound:",
                "`scripts/bench_armors.py` with IRMV_ARMOR_PROF=1 gives the cycles per ROI by phase (`profiles/r1_armors_phase_profile.json`).", ""]
    json.dump({"kernel": "conv_raster_kernel (Detect P3 box.0|cls.0, 3x3 64->128)", "frames": 128,
               "dram_bytes_per_launch": (float(h0.get("dram__bytes_read.sum", 0)) + float(h0.get("dram__bytes_write.sum", 0))) * 1e6,
               "gpu_time_us": float(h0.get("gpu__time_duration.sum", 0)),
               "tensor_pipe_pct": float(h0.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)),
               "source": "profiles/r1_raster_h0_raw.csv (ncu --set full)"},
              open(os.path.join(P, "r1_top_kernel.json"), "w"), indent=1)
    open(os.path.join(P, "r1_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
