"""profiles/r02_ncu_kernels.md + profiles/r02_ncu_traffic.json from the `ncu --set full` captures in gpurun_out/ (read with
`ncu -i ... --page raw --csv`; no GPU needed), and profiles/r02_launches_bf16.md from the launch list.
    python tools/make_profiles.py [molecules per launch, default 4144]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = int(sys.argv[1]) if len(sys.argv) > 1 else 4144
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
        ("launch__registers_per_thread", "registers / thread"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg-throttle / issue")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            d[h.split("TriageCompute.")[-1]] = (v, u)
        res.append(d)
    return res


def main():
    md = ["# Round 2 -- ncu --set full summaries of the hot tcgen05 kernels (BF16 mode, bench shape)\n",
          "Captured with `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 4 -c 2 python tools/prof_pair_train.py 4144` "
          "(a `PairTrainer` step over one 4144-pair micro-batch at H = 128, T = 6, N = 64, O = 128, head 8, K = 86; the same command had exited 0 "
          "without ncu first) on a B200. The `.ncu-rep` files stay in `gpurun_out/` (scratch). Durations under ncu are serialised, cold-cache "
          "numbers: compare shares and bytes, the timed figures are `bench.py`'s.\n"]
    traffic = {}
    for kern in ("ggnn_tc_bwd_kernel", "ggnn_tc_kernel", "wgrad2_kernel", "coattn_tc_kernel"):
        rep = os.path.join(ROOT, "gpurun_out", "r02_%s.ncu-rep" % kern)
        if not os.path.exists(rep):
            continue
        launches = rows_of(rep)
        for d in launches:
            name = d["Kernel Name"][0]
            md.append("### `%s`\n\n| metric | value | unit |\n|---|---|---|" % name)
            for k, label in KEYS:
                if k in d:
                    md.append("| %s | %s | %s |" % (label, d[k][0], d[k][1]))
            md.append("")
        def dram(d):
            return to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
        if kern == "wgrad2_kernel":       # the two launches of one encoder backward (stateless run + stateful run) together
            traffic["wgrad2_kernel"] = dict(dram_bytes=sum(dram(d) for d in launches), units=UNITS,
                                            note="sum over the launches of one encoder backward (one per run of equal statefulness)")
        elif kern == "coattn_tc_kernel":
            for d in launches:
                bwd = "true" in d["Kernel Name"][0] or ", 1," in d["Kernel Name"][0] or "(bool)1" in d["Kernel Name"][0]
                traffic["coattn_tc_kernel<128,%d,8>" % (1 if bwd else 0)] = dict(dram_bytes=dram(d), units=UNITS)
        else:
            key = "ggnn_tc_bwd_kernel<128,1>" if kern == "ggnn_tc_bwd_kernel" else "ggnn_tc_kernel<128,1,1>"
            traffic[key] = dict(dram_bytes=sum(dram(d) for d in launches) / len(launches), units=UNITS)
    open(os.path.join(ROOT, "profiles", "r02_ncu_kernels.md"), "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
