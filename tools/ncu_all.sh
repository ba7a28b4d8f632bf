#!/bin/bash
# Round evidence in one GPU call: ncu --set full captures of every kernel of the step at the BASELINE configs, and
# metric-only launch lists of whole steps. Run on the GPU box from the repo root (gpurun); reports land in gpurun_out/.
#   bash tools/ncu_all.sh r02
tag=${1:-r02}
out=gpurun_out
full="ncu --set full --clock-control none --import-source on"
cap() {   # name, kernel regex, skip, count, ncu_step args...
  name=$1; re=$2; skip=$3; cnt=$4; shift 4
  $full -k regex:"$re" --launch-skip $skip -c $cnt -f -o $out/ncu_${name}_$tag python tools/ncu_step.py "$@" > $out/ncu_${name}_$tag.log 2>&1
  ncu -i $out/ncu_${name}_$tag.ncu-rep --page raw --csv > $out/ncu_${name}_${tag}_raw.csv 2>/dev/null
  echo "$name: $(grep -c . $out/ncu_${name}_${tag}_raw.csv) csv lines"
}
cap k1_c3 'slab_reduce' 1 1 c3 kji 0 2
cap k1_c4 'slab_reduce' 1 1 c4 kji 0 2
cap k1_c5 'slab_reduce' 1 1 c5 kji 2048 2
cap k1_c2 'slab_reduce' 1 1 c2 kji 0 2
cap k1_ijk_c3 'slab_reduce' 1 1 c3 ijk 0 2
cap k23_c3 'gcm_to_les|les_to_gcm' 2 2 c3 kji 0 3
cap k23_c5 'gcm_to_les|les_to_gcm' 2 2 c5 kji 2048 3
cap k23_c3_256 'gcm_to_les|les_to_gcm' 2 2 c3 kji 256 3
cap k23_ijk_c3 'gcm_to_les|les_to_gcm|cloud_project' 3 3 c3 ijk 0 3
for cfg in "c3 kji 0" "c5 kji 2048" "c3 ijk 0" "c3 kji 256" "c2 kji 0" "c4 kji 0"; do
  set -- $cfg
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_$1_$2_$3_$tag.csv python tools/ncu_step.py $1 $2 $3 4 > /dev/null 2>&1
  echo "launch list $cfg: $(grep -c slab_reduce $out/launches_$1_$2_$3_$tag.csv) K1 launches"
done
ls -la $out/*_$tag.ncu-rep | awk '{print $5, $9}'
