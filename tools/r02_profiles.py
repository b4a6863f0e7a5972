#!/usr/bin/env python3
"""Regenerate the round-2 evidence under profiles/ from what the gpurun calls left in gpurun_out/ (run here, after a call).

    tools/r02_profiles.py

Copies the unprofiled bench lines (one JSON object per file), summarises the ncu reports of K1x (walk, raster) and K2 with
tools/ncu_summary.py / tools/ncu_lines.py, and writes profiles/r02_results.md.  Every source is optional: what is missing is
reported and skipped."""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")

# profile name -> candidate sources in gpurun_out (first that exists wins)
BENCH = {
    "r02_bench_n1.json": ["r2o_bench_n1.json", "r2h_bench_n1.json", "r2g_n1.json"],
    "r02_bench_reference_arm.json": ["r2o_reference.json", "r2h_reference.json", "r2b_reference.json"],
    "r02_bench_errors1pct.json": ["r2o_errors1pct.json", "r2h_errors1pct.json", "r2g_k2new_errors1pct.json"],
    "r02_bench_noise1pct_9472frames.json": ["r2o_noise.json", "r2i_noise.json", "r2g_noise.json"],
    "r02_bench_noise1pct_3552frames.json": ["r2o_noise_3552.json", "r2i_noise_3552.json", "r2d_noise_3552.json"],
    "r02_bench_color_correction_1.json": ["r2o_cc1.json", "r2h_cc1.json"],
    "r02_bench_color_correction_2.json": ["r2o_cc2.json", "r2h_cc2.json"],
    "r02_bench_mode4.json": ["r2s_mode4.json", "r2o_mode4.json", "r2h_mode4.json", "r2d_mode4.json"],
    "r02_bench_mode4_errors1pct.json": ["r2s_mode4_errors.json", "r2o_mode4_errors.json", "r2h_mode4_errors.json", "r2b_mode4_errors.json"],
    "r02_bench_mode67.json": ["r2o_mode67.json", "r2h_mode67.json"],
    "r02_bench_fountain_n1.json": ["r2o_fountain_n1.json", "r2h_fountain_n1.json", "r2g_fountain_n1.json"],
    "r02_bench_fountain_n2_window.json": ["r2p_fountain_n2.json", "r2g_fountain_n2.json"],
    "r02_bench_fountain_n2_nccl.json": ["r2j_fountain_n2_nccl.json", "r2g_fountain_n2_nccl.json"],
    "r02_bench_fountain_n8.json": ["r2q_fountain_n8.json", "r2k_fountain_n8.json"],
    "r02_bench_n2_window.json": ["r2p_n2_window.json"],
    "r02_bench_n2_window_direct.json": ["r2p_n2_window-direct.json", "r2g_n2_window.json"],
    "r02_bench_n2_nccl.json": ["r2p_n2_nccl.json", "r2g_n2_nccl.json"],
    "r02_bench_n2_torch_gather.json": ["r2g_n2_torch.json"],
    "r02_bench_n4_window_direct.json": ["r2k_n4_window.json"],
    "r02_bench_n4_window.json": ["r2r_n4_window.json"],
    "r02_bench_n1_second_run_with_walk_latency.json": ["r2r_n1.json"],
    "r02_bench_mode4_k2_per_warp_kernel.json": ["r2s_mode4_k2old.json"],
    "r02_bench_n8_window.json": ["r2q_n8_window.json"],
    "r02_bench_n8_window_direct.json": ["r2q_n8_window_direct.json", "r2k_n8_window.json"],
    "r02_bench_n8_nccl.json": ["r2k_n8_nccl.json"],
    "r02_bench_k2_per_warp_kernel_clean.json": ["r2g_k2old_clean.json"],
    "r02_bench_k2_per_warp_kernel_errors1pct.json": ["r2g_k2old_errors1pct.json"],
}


def last_json(path):
    for line in reversed(open(path).read().strip().split("\n")):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise ValueError("no JSON line in " + path)


def copy_bench():
    got = {}
    for name, cands in BENCH.items():
        for c in cands:
            p = os.path.join(GO, c)
            if os.path.exists(p):
                try:
                    j = last_json(p)
                except Exception as e:
                    print("skip %s: %s" % (c, e))
                    continue
                json.dump(j, open(os.path.join(PR, name), "w"))
                open(os.path.join(PR, name), "a").write("\n")
                got[name] = (c, j)
                break
        else:
            print("missing: %s (%s)" % (name, ", ".join(cands)))
    return got


def ncu_summary(rep, frames):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(frames)], capture_output=True, text=True).stdout
    return out.partition("TRAFFIC_JSON")[0].strip()


def ncu_lines(rep, pat, src, n):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, pat, os.path.join(ROOT, src), str(n)],
                          capture_output=True, text=True).stdout.strip()


def km(j):
    return j.get("kernel_ms_per_step", {})


def row(label, j, extra=""):
    k = km(j)
    e2e = (j.get("e2e") or {}).get("value")
    return "| %s | %s | %.3f | %s | %s | %s | %s | %s |" % (
        label, "{:,.0f}".format(j["value"]), j["ms_per_step"],
        "%.3f" % k["k1_decode"] if "k1_decode" in k else "", "%.3f" % k["k1x_flood_check"] if "k1x_flood_check" in k else "",
        "%.3f" % k["rs"] if "rs" in k else "", "{:,.0f}".format(e2e) if e2e else "", extra)


def main():
    os.makedirs(PR, exist_ok=True)
    got = copy_bench()
    md = ["# Round 2 -- measured results (B200, unprofiled `bench.py` lines; the JSON files next to this one are the lines themselves)", ""]
    md += ["| run | frames/s (device-resident) | ms/step | K1 ms | K1x ms | RS ms | e2e frames/s (host buffers) | note |", "|---|---|---|---|---|---|---|---|"]
    order = [("r02_bench_n1.json", "configs[1] clean, mode B, 10 000 frames/step"), ("r02_bench_errors1pct.json", "configs[2] 1 % wrong tiles"),
             ("r02_bench_color_correction_1.json", "colour correction 1"), ("r02_bench_color_correction_2.json", "colour correction 2 (reference default)"),
             ("r02_bench_mode4.json", "configs[4] mode 4C (legacy, RS(155,115)... see file)"), ("r02_bench_mode4_errors1pct.json", "mode 4C, 1 % wrong tiles"),
             ("r02_bench_mode4_k2_per_warp_kernel.json", "A/B: mode 4C with CB200_K2_FRAMES=0"),
             ("r02_bench_mode67.json", "mode Bm (67)"),
             ("r02_bench_noise1pct_9472frames.json", "1 % noise tiles: every frame through the exact walk, 9 472 frames"),
             ("r02_bench_noise1pct_3552frames.json", "same, 3 552 frames"),
             ("r02_bench_k2_per_warp_kernel_clean.json", "A/B: CB200_K2_FRAMES=0 (round-1-style per-warp RS kernel), clean"),
             ("r02_bench_k2_per_warp_kernel_errors1pct.json", "A/B: CB200_K2_FRAMES=0, 1 % wrong tiles"),
             ("r02_bench_n2_window.json", "N = 2, window, copy-engine push (default)"), ("r02_bench_n2_window_direct.json", "N = 2, window, direct stores of the RS kernel"),
             ("r02_bench_n2_nccl.json", "N = 2, cb200_gather_chunks (NCCL)"),
             ("r02_bench_n2_torch_gather.json", "N = 2, torch.distributed.gather (round-1 path)"),
             ("r02_bench_n4_window.json", "N = 4, window, copy-engine push (default)"), ("r02_bench_n4_window_direct.json", "N = 4, window, direct stores"),
             ("r02_bench_n8_window.json", "N = 8, window, copy-engine push (default)"), ("r02_bench_n8_window_direct.json", "N = 8, window, direct stores"),
             ("r02_bench_n8_nccl.json", "N = 8, NCCL")]
    for name, label in order:
        if name in got:
            src, j = got[name]
            md.append(row(label, j, "`%s`" % name))
    ref = got.get("r02_bench_reference_arm.json")
    if ref:
        j = ref[1]
        cb = j.get("cpu_baseline", {})
        md += ["", "Reference arm (`bench.py --impl reference`, `%s`): **%s frames/s** on %s threads (%s); stages per frame on one thread: %s" % (
            "r02_bench_reference_arm.json", "{:,.0f}".format(j["value"]), cb.get("cores"), cb.get("build", ""), json.dumps(cb.get("stages", {})))]
    lat = got.get("r02_bench_n1_second_run_with_walk_latency.json")
    if lat and lat[1].get("e2e_single_frame"):
        e = lat[1]["e2e_single_frame"]
        md += ["", "One frame per call (`e2e_single_frame`, `r02_bench_n1_second_run_with_walk_latency.json`): clean frame %.3f ms median from pinned memory "
               "(%.3f ms from pageable); a frame that needs the exact walk: **%.1f ms** -- the walk is one warp per frame, its throughput comes from "
               "thousands of frames in flight." % (e["pinned"]["median_ms"], e["pageable"]["median_ms"], (e.get("exact_walk_frame") or {}).get("median_ms", float("nan")))]
    for name in ("r02_bench_fountain_n1.json", "r02_bench_fountain_n2_window.json", "r02_bench_fountain_n2_nccl.json", "r02_bench_fountain_n8.json"):
        if name in got:
            j = got[name][1]
            md += ["", "Fountain (configs[3], `%s`): %s; device %.2f ms per transfer (%s frames/s), sink %s chunks/s, end to end %.0f MB/s of file. %s" % (
                name, j.get("parity"), j["ms_per_step"], "{:,.0f}".format(j["value"]), "{:,.0f}".format(j["rank0_sink"]["chunks_per_s"]),
                j["end_to_end"]["file_MB_per_s"], j.get("saturation", ""))]
    open(os.path.join(PR, "r02_results.md"), "w").write("\n".join(md) + "\n")
    print("wrote profiles/r02_results.md with %d bench lines" % len(got))
    ncu_docs()


def first(*names):
    for n in names:
        p = os.path.join(GO, n)
        if os.path.exists(p):
            return p
    return None


def launch_lists():
    import collections, csv, re
    out = []
    for src, dst, title in (("r2o_launches.csv", "r02_launch_list_ncu.csv", "bench.py --steps 2 --warmup 1 (clean frames)"),
                            ("r2o_launches_noise.csv", "r02_launch_list_k1x_noise1pct.csv", "bench.py --workload noise1pct --frames 4736 --steps 1 --warmup 1"),
                            ("r2i_launches.csv", "r02_launch_list_ncu.csv", "bench.py --steps 2 --warmup 1 (clean frames)"),
                            ("r2i_launches_noise.csv", "r02_launch_list_k1x_noise1pct.csv", "bench.py --workload noise1pct --frames 4736 --steps 1 --warmup 1")):
        p = os.path.join(GO, src)
        if not os.path.exists(p) or any(dst in o for o in out):
            continue
        rows = [l for l in open(p) if l.startswith('"')]
        open(os.path.join(PR, dst), "w").write("".join(rows))
        rd = list(csv.reader(rows))
        hdr, rd = rd[0], rd[1:]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = collections.OrderedDict()
        for r in rd:
            n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("cb200::", "")
            agg.setdefault(n, []).append(float(r[vi].replace(",", "")) / 1e6)
        dec = [k for k in agg if k.startswith(("k1_decode", "k_flood", "k_rs_", "k_chunk_mask", "k_pack"))]
        if not dec:
            continue
        steps = max(len(agg[k]) for k in dec)
        tot = sum(sum(agg[k]) for k in dec) / steps
        tab = "### `%s` -- %s\n\n| kernel | launches | avg ms (cold, serialised under ncu) | share of the decode step |\n|---|---|---|---|\n" % (dst, title)
        for k in dec:
            a = sum(agg[k]) / steps
            tab += "| %s | %d | %.3f | %.1f %% |\n" % (k, len(agg[k]), a, 100 * a / tot)
        out.append(tab)
    return "\n".join(out)


def sass_evidence():
    lib = os.path.join(ROOT, "libcimbar_b200", "lib", "libcb200.so")
    if not os.path.exists(lib):
        return ""
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    import collections, re
    cur, per = None, collections.OrderedDict()
    pats = ["UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "IDP.2A", "IDP.4A", "IMAD.HI", "REDUX", "POPC", "PRMT", "SHFL", "MATCH", "VOTE", "BAR.SYNC", "STL", "LDL",
            "HMMA", "UTMALDG", "UTCMMA"]
    for line in txt.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur:
            for p in pats:
                if re.search(r"\b" + re.escape(p), line):
                    per[cur][p] += 1
    want = ["k1_decode_kernelILi4ELb1ELi0", "k_rs_frames", "k_rs_decodeILi1ELb1", "k_flood_walk", "k_flood_raster_fast", "k_deskew", "k_gather_publish", "k_gather_wait"]
    out = ["## SASS evidence (`cuobjdump -sass libcimbar_b200/lib/libcb200.so`, instruction counts per kernel)", "",
           "| kernel | " + " | ".join(pats) + " |", "|---|" + "---|" * len(pats)]
    for fn, c in per.items():
        if any(w in fn for w in want):
            short = re.sub(r"^_ZN5cb200\d+", "", fn)[:44]
            out.append("| `%s` | " % short + " | ".join(str(c.get(p, 0)) for p in pats) + " |")
    out += ["", "`UBLKCP` = `cp.async.bulk` (TMA bulk copy, K1's stage loads); `SYNCS` = mbarrier operations; `LDGSTS` = `cp.async` (K2's cell-byte",
            "staging); `IDP.2A/4A` = integer dot products (gray conversion, colour sums); `IMAD.HI` = the bit gather of the threshold",
            "bits; `REDUX` = warp reductions (heap pop, Berlekamp-Massey, Forney). No tensor-core instruction (`HMMA`/`UTCMMA`) and no",
            "tensor-map TMA (`UTMALDG`): the path is byte/bit work whose loads are contiguous rows."]
    return "\n".join(out)


def ncu_docs():
    """K1 / K2 / K1x summaries of the `ncu --set full` captures (numbers under ncu are never bench values)"""
    k1 = first("r2o_k1.ncu-rep", "r2i_k1.ncu-rep")
    k2 = first("r2o_k2.ncu-rep", "r2i_k2.ncu-rep", "r2h_k2.ncu-rep")
    k2e = first("r2o_k2_err.ncu-rep", "r2h_k2_err.ncu-rep")
    walk = first("r2o_walk.ncu-rep", "r2d_walk.ncu-rep")
    rast = first("r2o_raster.ncu-rep", "r2a_raster.ncu-rep")
    out = ["# Round 2 -- ncu evidence (B200, `ncu --set full --clock-control none --import-source on`, one launch each;",
           "summaries by `tools/ncu_summary.py` / `tools/ncu_lines.py`; the unprofiled numbers are in `r02_results.md`)", ""]
    if k1:
        body = ncu_summary(k1, 4144)
        out += ["## K1 `k1_decode_kernel<4,true,0>` (4 144 frames in the launch) -- `%s`" % os.path.basename(k1), body, "", "```",
                ncu_lines(k1, "k1_decode:k1_decode_kernelILi4ELb1E", "libcimbar_b200/csrc/k1_decode.cu", 14), "```", ""]
        tj = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), k1, "4144"], capture_output=True, text=True).stdout.partition("TRAFFIC_JSON")[2].strip()
        if tj:
            t = json.loads(tj)
            t["source"] = ("profiles/r02_ncu_summary.md (ncu --set full --clock-control none, k1_decode_kernel<4,true,0>, 4144 frames in the "
                           "launch: dram__bytes_read.sum + dram__bytes_write.sum)")
            json.dump(t, open(os.path.join(PR, "k1_traffic.json"), "w"), indent=1)
    if k2:
        out += ["## K2 `k_rs_frames`, clean frames (4 144 in the launch) -- `%s`" % os.path.basename(k2), ncu_summary(k2, 4144), "", "```",
                ncu_lines(k2, "k_rs_frames", "libcimbar_b200/csrc/k2_rs.cu", 16), "```", ""]
    if k2e:
        out += ["## K2 `k_rs_frames`, 1 %% wrong tiles (every block is corrected) -- `%s`" % os.path.basename(k2e), ncu_summary(k2e, 4144), "", "```",
                ncu_lines(k2e, "k_rs_frames", "libcimbar_b200/csrc/k2_rs.cu", 16), "```", ""]
    ll = launch_lists()
    if ll:
        out += ["## Launch lists (`ncu --metrics gpu__time_duration.sum --clock-control none`: every launch of the command)", "", ll, ""]
    out += [sass_evidence(), ""]
    open(os.path.join(PR, "r02_ncu_summary.md"), "w").write("\n".join(out) + "\n")
    w = []
    if walk:
        w += ["## `k_flood_walk` (4 736 frames = one wave of 32 walks per SM) -- `%s`" % os.path.basename(walk), ncu_summary(walk, 4736), "", "```",
              ncu_lines(walk, "k_flood_walk", "libcimbar_b200/csrc/k1x_flood.cu", 24), "```", ""]
    if rast:
        w += ["## `k_flood_raster_fast` -- `%s`" % os.path.basename(rast), ncu_summary(rast, 1776), ""]
    head_p = os.path.join(PR, "r02_k1x_walk_head.md")
    head = open(head_p).read() if os.path.exists(head_p) else "# Round 2 -- K1x (exact flood walk)\n"
    open(os.path.join(PR, "r02_k1x_walk.md"), "w").write(head + "\n" + "\n".join(w) + "\n")
    print("wrote profiles/r02_ncu_summary.md, profiles/r02_k1x_walk.md")


if __name__ == "__main__":
    main()
