"""Copy the evidence of tools/round_evidence.sh (gpurun_out/ev/, scratch) into profiles/ (tracked):

    python tools/collect_profiles.py [round]        # default round = 2  ->  profiles/r2_*

bench lines (last JSON line of each run), phase timelines, the GPU test log, and the ncu launch lists reduced to
(id, kernel, microseconds).  ncu --set full captures are summarised separately by tools/ncu_summary.py."""
import csv
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E = os.path.join(ROOT, "gpurun_out", "ev")
P = os.path.join(ROOT, "profiles")


def main():
    r = "r%s_" % (sys.argv[1] if len(sys.argv) > 1 else "2")
    names = {"bench_c2.json": r + "bench_ours.json", "bench_c2_reference.json": r + "bench_ref.json"}
    done = []
    for f in sorted(os.listdir(E)):
        src = os.path.join(E, f)
        if f.startswith("bench_") and f.endswith(".json"):
            lines = open(src).read().strip().splitlines()
            if lines and lines[-1].startswith("{"):
                dst = names.get(f, r + f)
                open(os.path.join(P, dst), "w").write(lines[-1] + "\n")
                done.append(dst)
        elif f.startswith("timeline_") and f.endswith(".json"):
            shutil.copy(src, os.path.join(P, r + f))
            done.append(r + f)
        elif f == "tests.log":
            shutil.copy(src, os.path.join(P, r + "gpu_tests.log"))
            done.append(r + "gpu_tests.log")
        elif f == "tests_quick.log":                # QUICK=1 pass on the final tree (default paths + fast LSE + smoke)
            shutil.copy(src, os.path.join(P, r + "gpu_tests_final.log"))
            done.append(r + "gpu_tests_final.log")
        elif f.startswith("launches_") and f.endswith(".csv"):
            rows = [x for x in csv.reader(open(src)) if len(x) > 5]
            if not rows:
                continue
            h = rows[0]
            ki, vi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
            with open(os.path.join(P, r + f), "w") as o:
                o.write("id,kernel,gpu__time_duration_us\n")
                for x in rows[1:]:
                    o.write('%s,"%s",%.2f\n' % (x[ii], x[ki][:90].replace('"', "'"), float(x[vi].replace(",", "")) / 1e3))
            done.append(r + f)
    print("\n".join(done))


if __name__ == "__main__":
    main()
