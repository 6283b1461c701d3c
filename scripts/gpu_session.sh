# round 2, session o: where does the time of a 1/8-share frame go with the tail hand-off on?  per-kernel durations and instruction counts
set -x
cd $GRAFT_REPO_ROOT
for a in 0 4 16; do
OCLR_HANDOFF_AFTER=$a ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k "regex:wf_pipe|wf_tail" -c 12 --csv --log-file gpurun_out/r02o_after$a.csv python scripts/ncu_target_band.py 2 8 2 > gpurun_out/r02o_after$a.log 2>&1
done
OCLR_HANDOFF_MAX_PATHS=0 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k "regex:wf_pipe|wf_tail" -c 12 --csv --log-file gpurun_out/r02o_off.csv python scripts/ncu_target_band.py 2 8 2 > gpurun_out/r02o_off.log 2>&1
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/r02o_*.csv")):
    rows = list(csv.reader(open(f)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ix = {k: i for i, k in enumerate(rows[h])}
    print(f)
    for r in rows[h + 1:]:
        if len(r) > ix["Metric Value"]:
            print("  ", r[ix["ID"]], r[ix["Kernel Name"]][:40], r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]])
PY
