#!/bin/bash
# ncu --set full of the first invocation of every kernel function of the whole path at 3.1 Gbp (one process).
# The report stays on the box (it is larger than what gpurun brings back); its summaries go to gpurun_out/.
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-id ::regex:.*:1 -f -o /tmp/r02_full_all \
    python profiles/scripts/r02_ncu_target.py all > gpurun_out/r02_full_all_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02_full_all_ncu.log
python profiles/scripts/summarize_ncu.py /tmp/r02_full_all.ncu-rep > gpurun_out/r02_ncu_full_3100mbp.md
ncu -i /tmp/r02_full_all.ncu-rep --page raw --csv > /tmp/raw.csv; gzip -c /tmp/raw.csv > gpurun_out/r02_ncu_full_3100mbp_raw.csv.gz
ls -la /tmp/r02_full_all.ncu-rep gpurun_out/r02_ncu_full_3100mbp*; wc -l gpurun_out/r02_ncu_full_3100mbp.md
