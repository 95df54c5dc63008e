"""CPU side of tools/gpu_profile_r02.sh: turns gpurun_out/*_<tag>.* into the tracked summaries under profiles/.

    python tools/collect_profiles_r02.py [tag]
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def have(name):
    return os.path.exists(os.path.join(OUT, name))


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[2:]


def main():
    for f in ("pytest_%s.log" % tag, "bench_%s_n1.json" % tag, "bench_%s_reference.json" % tag):
        if have(f):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f.replace("pytest_", "pytest_gpu_")))
    for name in ("1920", "1m", "131k", "bf_1m", "wd64_1920", "ref"):
        f = "launches_%s_%s.csv" % (name, tag)
        if have(f):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f))
    traffic = {"source": "ncu --set full --clock-control none (tools/gpu_profile_r02.sh %s); per launch, cold cache" % tag, "configs": {}}
    for name in ("1920", "1m", "131k", "bf_1m", "bf", "wd64_1920"):
        rep = os.path.join(OUT, "prof_%s_%s.ncu-rep" % (name, tag))
        if not os.path.exists(rep):
            continue
        with open(os.path.join(PROF, "ncu_%s_%s.txt" % (name, tag)), "w") as fh:
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), rep], stdout=fh, check=True)
        hdr, data = raw_rows(rep)
        col = {h: i for i, h in enumerate(hdr)}
        cfg = {}
        for r in data:
            kname = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]

            def num(key, scale=1.0):
                try:
                    return float(r[col[key]]) * scale
                except (KeyError, ValueError):
                    return None
            # ncu prints byte counts in the unit row's unit; ask for base units through --print-units base instead
            cfg[kname] = {"dram_read_bytes": None, "dram_write_bytes": None, "duration_us": None}
        base = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(base)))
        col = {h: i for i, h in enumerate(rows[0])}
        for r in rows[2:]:
            kname = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
            cfg[kname] = {"dram_read_bytes": float(r[col["dram__bytes_read.sum"]]), "dram_write_bytes": float(r[col["dram__bytes_write.sum"]]),
                          "duration_us": float(r[col["gpu__time_duration.sum"]]) / 1e3}
        traffic["configs"][name] = cfg
        kern = {"1920": "rollout_half", "1m": "rollout_tc", "131k": "rollout_tc", "bf_1m": "rollout_kernel", "bf": "rollout_bf", "wd64_1920": "rollout_pipe64"}[name]
        with open(os.path.join(PROF, "ncu_%s_%s_lines.txt" % (name, tag)), "w") as fh:
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kern, "40"], stdout=fh)
    if traffic["configs"]:
        json.dump(traffic, open(os.path.join(PROF, "ncu_traffic_%s.json" % tag), "w"), indent=1)
    print(json.dumps(traffic, indent=1)[:3000])


if __name__ == "__main__":
    main()
