#!/bin/bash
# Per-kernel counts of the SASS mnemonics that show what the shipped library is made of (bulk copies / mbarriers /
# fp64 FMAs / tensor-core ops): tools/sass_summary.sh > profiles/sass_summary.txt
lib=${1:-bout-dev_b200/csrc/libbout_b200.so}
echo "# $(basename $lib): cuobjdump -lelf:"; cuobjdump -lelf $lib | sed 's/^/#   /'
cuobjdump -sass $lib | awk '
/Function :/ {fn=$3}
/UBLKCP/ {ublkcp[fn]++} /UTMALDG|UTMASTG/ {utma[fn]++} /SYNCS/ {syncs[fn]++} /DFMA/ {dfma[fn]++} /DADD|DMUL/ {dadd[fn]++}
/LDGSTS/ {ldgsts[fn]++} /HMMA|UTCMMA|QGMMA/ {mma[fn]++} /STL|LDL/ {local[fn]++} /ACQBULK|UBLKPF/ {pf[fn]++}
END {printf "%-70s %7s %7s %7s %7s %8s %6s %6s %6s\n","kernel","UBLKCP","UBLKPF","SYNCS","DFMA","DADD/MUL","LDGSTS","MMA","LDL/STL";
 for (f in dfma) all[f]=1; for (f in ublkcp) all[f]=1; for (f in syncs) all[f]=1; for (f in dadd) all[f]=1;
 for (f in all) printf "%-70s %7d %7d %7d %7d %8d %6d %6d %6d\n", substr(f,1,70), ublkcp[f], pf[f], syncs[f], dfma[f], dadd[f], ldgsts[f], mma[f], local[f]}' | (read h; echo "$h"; sort)
