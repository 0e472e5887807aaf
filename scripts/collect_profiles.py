#!/usr/bin/env python
"""Turn the files a `scripts/gpu_round.sh <tag>` / `scripts/gpu_profile.sh <tag>_<wl> <wl>` call left in gpurun_out/ into the
committed evidence under profiles/: raw + per-line ncu summaries, launch lists, the bench line, kernel_counters.json.

    python scripts/collect_profiles.py <tag> [workload ...]
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TICKS = {"c2": 8192, "c3": 1024, "c4": 1024, "c5": 1024}
ENVS = {"c2": 4096, "c3": 20480, "c4": 8192, "c5": 8192}


def main():
    tag, wls = sys.argv[1], sys.argv[2:] or ["c4"]
    for name in (f"{tag}_bench.json", f"{tag}_gputest.log"):
        src = os.path.join(OUT, name)
        if os.path.exists(src):
            dst = name.replace("_bench.json", "_bench_default_1gpu.json").replace("_gputest.log", "_gputest.txt")
            shutil.copy(src, os.path.join(PROF, dst))
    kc_path = os.path.join(PROF, "kernel_counters.json")
    kc = json.load(open(kc_path)) if os.path.exists(kc_path) else {}
    for wl in wls:
        rep = os.path.join(OUT, f"{tag}_{wl}_prof.ncu-rep")
        raw = os.path.join(PROF, f"{tag}_{wl}_k_run_raw.csv")
        lines = os.path.join(PROF, f"{tag}_{wl}_k_run_lines.txt")
        open(raw, "w").write(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
        open(lines, "w").write(subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, "40"], capture_output=True, text=True).stdout)
        shutil.copy(os.path.join(OUT, f"{tag}_{wl}_launches.csv"), os.path.join(PROF, f"{tag}_{wl}_launches.csv"))
        rows = list(csv.reader(open(raw)))
        d = {h: (v, u) for h, v, u in zip(rows[0], rows[-1], rows[1])}
        f = lambda k: float(d[k][0].replace(",", ""))
        to_bytes = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        dram = sum(f(k) * to_bytes[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        plain = json.loads(open(os.path.join(OUT, f"{tag}_{wl}_plain.log")).read().strip().splitlines()[-1])
        dpl = plain["config"]["decisions_per_launch"]
        wi = f("smsp__inst_executed.sum")
        ms = f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[d["gpu__time_duration.sum"][1]]
        kc[wl] = {"ticks": TICKS[wl], "envs": ENVS[wl], "kernel": d["Kernel Name"][0], "dram_bytes_per_launch": dram, "warp_inst_per_launch": wi,
                  "decisions_per_launch": dpl, "warp_inst_per_decision": wi / dpl,
                  "threads_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"), "dram_bytes_per_decision": dram / dpl,
                  "kernel_ms_under_ncu": ms, "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"), "registers": int(f("launch__registers_per_thread")),
                  "l1_hit_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": f("lts__t_sector_hit_rate.pct"),
                  "source": f"profiles/{tag}_{wl}_k_run_raw.csv (ncu --set full, one warm k_run launch of `bench.py --workload {wl} --steps 2 --warmup 2`; "
                            f"decisions per launch from the plain run of the same command)"}
        print(wl, json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in kc[wl].items() if k != "source"}))
        print(open(lines).read().splitlines()[1])
    json.dump(kc, open(kc_path, "w"), indent=1)


if __name__ == "__main__":
    main()
