"""Summarise `ncu -i X.ncu-rep --page source --csv` : top SASS instructions by stall samples, per kernel."""
import csv
import subprocess
import sys


def main(rep, pattern, ntop=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pattern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernel, hdr, data = None, None, []
    blocks = []
    for r in rows:
        if r and r[0] == "Kernel Name":
            if hdr:
                blocks.append((kernel, hdr, data))
            kernel, hdr, data = r[1], None, []
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) == len(hdr):
            data.append(r)
    if hdr:
        blocks.append((kernel, hdr, data))
    for kernel, hdr, data in blocks[:1]:
        ci = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ci["# Samples"]]) for r in data)
        print(f"== {kernel[:90]}  total samples {tot}")
        for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:ntop]:
            n = int(r[ci["# Samples"]])
            st = {h[6:]: int(r[ci[h]]) for h in hdr
                  if h.startswith("stall_") and "Not Issued" not in h and r[ci[h]].isdigit() and int(r[ci[h]]) > 0}
            top3 = sorted(st.items(), key=lambda x: -x[1])[:3]
            print(f"{100.0 * n / tot:5.1f}%  exec={r[ci['Instructions Executed']]:>8}  {r[ci['Source']].strip()[:64]:64s} {top3}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
