#!/bin/bash
# Per-kernel SASS evidence (runs on the CPU box: cuobjdump only).  For every __global__ in the built objects writes
#   profiles/sass/<kernel>.sass.gz   full listing
# and one summary table profiles/sass/SUMMARY.md with instruction count and the Blackwell-specific mnemonics
# (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG/UBLKCP = TMA, LDGMC/STGMC = multimem, SYNCS = mbarrier).
set -e
OUT=profiles/sass
mkdir -p $OUT
python -m midgpt_b200.ops.build > /dev/null
echo "| kernel | SASS instr | UTC*MMA | LDTM | STTM | UTMALDG | UTMASTG | UTMAREDG | UBLKCP | SYNCS | LDGMC/STGMC/REDG | HMMA (legacy) |" > $OUT/SUMMARY.md
echo "|---|---|---|---|---|---|---|---|---|---|---|---|" >> $OUT/SUMMARY.md
for obj in build/obj/*.cu.o; do
  for fn in $(cuobjdump -sass $obj | grep -oE "Function : [A-Za-z0-9_]+" | awk '{print $3}'); do
    short=$(echo $fn | c++filt | sed -E 's/\(.*//; s/^void //; s/mg:://; s/[<>, ]+/_/g; s/_$//' | cut -c1-80)
    cuobjdump -sass -fun $fn $obj > /tmp/k.sass 2>/dev/null || continue
    gzip -c /tmp/k.sass > $OUT/$short.sass.gz
    n=$(grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass || true)
    c() { grep -cE "$1" /tmp/k.sass || true; }
    echo "| \`$short\` | $n | $(c 'UTC[A-Z]*MMA') | $(c 'LDTM') | $(c 'STTM') | $(c 'UTMALDG') | $(c 'UTMASTG') | $(c 'UTMAREDG') | $(c 'UBLKCP') | $(c 'SYNCS') | $(c 'LDGMC|STGMC|REDG') | $(c ' HMMA') |" >> $OUT/SUMMARY.md
  done
done
cat $OUT/SUMMARY.md | head -50
du -sh $OUT
