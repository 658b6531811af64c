"""Per-layer table from an ncu launch list (gpu__time_duration) of scripts/profile_replay.py."""
import csv
import re
import sys


def layers(split=False, fused=True):
    """GEMM launches of one replay in issue order (conv0 first; it runs inside the stem kernel).
    fused=True: the production program, where m2.cv1 and the Detect towers' final 1x1 convs run as
    fused tails of their producers (engine.cu fuse_tails)."""
    def c2f(name, c1, c2, n, hw, up=0):
        c = c2 // 2
        # neck C2f whose cv1 reads concat(upsample(a), b): split into W_a at half resolution + W_b (engine.cu)
        out = [(f"{name}.cv1a", hw // 2, up, 2 * c, 1, 1), (f"{name}.cv1b", hw, c1 - up, 2 * c, 1, 1)] if (up and split) else \
              [(f"{name}.cv1", hw, c1, 2 * c, 1, 1)]
        for i in range(n):
            out += [(f"{name}.m{i}.cv1", hw, c, c, 3, 1), (f"{name}.m{i}.cv2", hw, c, c, 3, 1)]
        out.append((f"{name}.cv2", hw, (2 + n) * c, c2, 1, 1))
        return out
    d = [("m0", 320, 8, 16, 3, 2), ("m1", 160, 16, 32, 3, 2)] + c2f("m2", 32, 32, 1, 160)
    d += [("m3", 80, 32, 64, 3, 2)] + c2f("m4", 64, 64, 2, 80) + [("m5", 40, 64, 128, 3, 2)] + c2f("m6", 128, 128, 2, 40)
    d += [("m7", 20, 128, 256, 3, 2)] + c2f("m8", 256, 256, 1, 20) + [("m9.cv1", 20, 256, 128, 1, 1), ("POOL", 20, 0, 0, 0, 0), ("m9.cv2", 20, 512, 256, 1, 1)]
    d += c2f("m12", 384, 128, 1, 40, up=256) + c2f("m15", 192, 64, 1, 80, up=128) + [("m16", 40, 64, 64, 3, 2)] + c2f("m18", 192, 128, 1, 40)
    d += [("m19", 20, 128, 128, 3, 2)] + c2f("m21", 384, 256, 1, 20)
    for i, (hw, ch) in enumerate(((80, 64), (40, 128), (20, 256))):
        d += [(f"h{i}.01", hw, ch, 128, 3, 1), (f"h{i}.box1", hw, 64, 64, 3, 1), (f"h{i}.box2", hw, 64, 64, 1, 1),
              (f"h{i}.cls1", hw, 64, 64, 3, 1), (f"h{i}.cls2", hw, 64, 16, 1, 1)]
    if fused:
        drop = {"m2.cv1": "m1", "m4.cv1": "m3"}
        for i in range(3):
            drop[f"h{i}.box2"] = f"h{i}.box1"
            drop[f"h{i}.cls2"] = f"h{i}.cls1"
        d = [((n + "+1x1") if n in drop.values() else n, hw, cin, cout, k, s) for (n, hw, cin, cout, k, s) in d if n not in drop]
    return d


def name_ops(ops):
    """Names for the engine's launches (YoloEngine.describe_ops()): walks the unfused layer list, a
    launch with a fused 1x1 tail consumes two entries.  Returns (name, hw, cin, cout, k, s, flop_per_frame)."""
    L = layers(fused=False)[1:]
    out, i = [], 0
    split_b = None
    for o in ops:
        if o["kind"] == "pool":
            assert L[i][0] == "POOL"
            out.append(("POOL", 20, 0, 0, 0, 0, 0.0))
            i += 1
            continue
        if split_b:                       # second launch of a 1x1 over concat(upsample(a), b): W_b at full resolution
            n, hw, cin, cout, k, s = split_b
            split_b = None
            out.append((n + "b", hw, o["cin"], cout, k, s, 2.0 * hw * hw * o["cin"] * cout))
            continue
        n, hw, cin, cout, k, s = L[i]
        i += 1
        if n in ("m12.cv1", "m15.cv1") and o["cin"] != cin:
            # first launch: W_a at the upsampled input's own (half) resolution, no bias / activation
            split_b = (n, hw, cin, cout, k, s)
            out.append((n + "a", hw // 2, o["cin"], cout, k, s, 2.0 * (hw // 2) ** 2 * o["cin"] * cout))
            continue
        fl = 2.0 * hw * hw * k * k * cin * cout
        if o["tail_cout"]:
            n2, hw2, cin2, cout2, k2, s2 = L[i]
            fl += 2.0 * hw2 * hw2 * cin2 * cout2
            n = f"{n}+{n2.split('.')[-1]}"
            i += 1
        out.append((n, hw, cin, cout, k, s, fl))
    assert i == len(L), (i, len(L))
    return out


def main():
    path, B = sys.argv[1], int(sys.argv[2])
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"])) for r in rows]
    idx = [i for i, n in enumerate(names) if "preprocess_kernel" in n[0]]
    seq = names[idx[-1]:idx[-1] + 66]
    out, tot = [], 0.0
    for (n, v), d in zip(seq[1:62], layers()):
        name, hw, cin, cout, k, s = d
        kern = re.sub(r"\(.*", "", n).split("::")[-1].replace("_kernel", "")
        if name == "POOL":
            out.append((name, kern, v, 0.0, 0.0))
            tot += v
            continue
        M = B * hw * hw
        fl = 2.0 * M * k * k * cin * cout
        io = (B * (hw * s) ** 2 * cin + M * cout) * 2.0
        out.append((name, kern, v, fl / v / 1e3, io / v))
        tot += v
    print(f"conv+pool total {tot/1e3:.1f} us for {B} frames = {tot/1e3/B:.2f} us/frame; other:",
          [(re.sub(r'\(.*', '', n).split('::')[-1], round(v / 1e3, 1)) for n, v in [seq[0]] + seq[62:]])
    print(f"{'layer':12s} {'kernel':12s} {'us':>8s} {'share':>6s} {'TFLOP/s':>8s} {'GB/s(in+out)':>12s}")
    for name, kern, v, tf, gb in sorted(out, key=lambda r: -r[2]):
        print(f"{name:12s} {kern:12s} {v/1e3:8.1f} {100*v/tot:5.1f}% {tf:8.1f} {gb:12.1f}")


if __name__ == "__main__":
    main()
