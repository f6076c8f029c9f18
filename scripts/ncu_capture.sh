# usage: ncu_capture.sh TAG KERNEL_REGEX SKIP -- command...   (run on the GPU box)
# One `ncu --set full` capture of launch number SKIP of the kernels matching KERNEL_REGEX, after the same command has
# exited 0 without ncu.  Leaves gpurun_out/TAG.summary.txt (scripts/ncu_summary.py), TAG.raw.csv.gz and
# TAG.source.csv.gz; the .ncu-rep itself is deleted (gpurun only brings back 64 MiB).
TAG=$1; RX=$2; SKIP=$3; shift 4
O=gpurun_out
timeout 200 "$@" > $O/$TAG.plain.log 2>&1 || { echo "plain run failed for $TAG"; tail -5 $O/$TAG.plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -f -o $O/$TAG "$@" > $O/$TAG.ncu.log 2>&1
echo "ncu $TAG rc=$?"
python scripts/ncu_summary.py $O/$TAG.ncu-rep > $O/$TAG.summary.txt 2>&1
ncu -i $O/$TAG.ncu-rep --page raw --csv 2>/dev/null | gzip > $O/$TAG.raw.csv.gz
ncu -i $O/$TAG.ncu-rep --page source --csv 2>/dev/null | gzip > $O/$TAG.source.csv.gz
rm -f $O/$TAG.ncu-rep
head -40 $O/$TAG.summary.txt
