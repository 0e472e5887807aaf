#!/usr/bin/env python
"""Summarise an ncu report per SOURCE LINE: warp instructions executed, stall samples and the dominant stall reason.

    python scripts/ncu_lines.py gpurun_out/prof_X.ncu-rep [top_n] > profiles/rNN_X_lines.txt

Reads `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo and --import-source on)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            def num(k):
                try:
                    return float(d.get(k, "0").replace(",", ""))
                except ValueError:
                    return 0.0
            stalls = {k[6:]: num(k) for k in d if k.startswith("stall_") and "Not Issued" not in k}
            lines.append((cur_file, int(r[0]), r[1].strip(), num("Instructions Executed"), num("# Samples"),
                          num("Thread Instructions Executed"), stalls))
    tot_i = sum(x[3] for x in lines) or 1.0
    tot_s = sum(x[4] for x in lines) or 1.0
    tot_t = sum(x[5] for x in lines)
    print(f"total warp instructions {tot_i:.4g}, thread instructions {tot_t:.4g} (avg {tot_t / tot_i:.2f} active threads/inst), "
          f"stall samples {tot_s:.4g}")
    agg = {}
    for x in lines:
        for k, v in x[6].items():
            agg[k] = agg.get(k, 0.0) + v
    print("stall reasons (all samples): " + ", ".join(f"{k} {100 * v / tot_s:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print(f"\n--- top {top} lines by stall samples")
    print(f"{'file:line':<22}{'inst%':>7}{'smpl%':>7}{'thr/inst':>9}  {'top stalls':<34} source")
    for x in sorted(lines, key=lambda x: -x[4])[:top]:
        st = sorted(x[6].items(), key=lambda kv: -kv[1])[:2]
        sts = " ".join(f"{k}:{100 * v / max(x[4], 1):.0f}%" for k, v in st if v)
        print(f"{x[0] + ':' + str(x[1]):<22}{100 * x[3] / tot_i:>7.2f}{100 * x[4] / tot_s:>7.2f}{x[5] / max(x[3], 1):>9.1f}  {sts:<34} {x[2][:90]}")
    print(f"\n--- top {top} lines by warp instructions executed")
    for x in sorted(lines, key=lambda x: -x[3])[:top]:
        print(f"{x[0] + ':' + str(x[1]):<22}{100 * x[3] / tot_i:>7.2f}{100 * x[4] / tot_s:>7.2f}{x[5] / max(x[3], 1):>9.1f}  {x[2][:110]}")


if __name__ == "__main__":
    main()
