#!/bin/bash
# A/B of the light kernels (control-step prologue / epilogue, state provider) between library builds:
#   tools/gpu/ab_light.sh <tag> <lib1> <lib2> ...   (lib = "default" or a name under ab/ without .so)
# per-launch kernel durations come from an ncu launch list (cold-cache, serialised: compare like with like);
# the event-timed control step of tools/perf_light.py is printed beside them.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
tag=$1; shift
for lib in "$@"; do
  if [ $lib = default ]; then unset RG_CUDA_LIB; else export RG_CUDA_LIB=$PWD/ab/$lib.so; fi
  out=gpurun_out/${tag}_light_$lib
  python tools/perf_light.py ${AB_LIGHT_N:-1048576} > $out.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"step_|state_from_sim" --csv --log-file $out.csv python tools/perf_light.py ${AB_LIGHT_N:-1048576} > /dev/null 2>&1
  echo "== $lib"; cat $out.log
  python - "$out.csv" <<'PY'
import csv, collections, statistics, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) > mv:
        v = float(r[mv].replace(",", "")); us = v / 1e3 if r[mu] in ("ns", "nsecond") else v
        agg.setdefault(r[kn].split("(")[0].split("::")[-1], []).append(us)
for k, v in agg.items(): print(f"   {k}: median {statistics.median(v):.1f} us over {len(v)} launches")
PY
done
