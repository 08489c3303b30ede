#!/bin/bash
# profiles/sass_rN.md: which Blackwell-native instructions the in-tree library really contains.
SO=tf2_gnn_b200/csrc/libtfgnn_b200.so
OUT=${1:-profiles/sass_r2.md}
cuobjdump -sass $SO > /tmp/tfgnn_sass.txt
{
  echo "# SASS evidence: \`cuobjdump -sass $SO\` at $(git log --oneline | head -1 | cut -c1-7)"
  echo
  echo "| SASS mnemonic | what it is | count |"
  echo "|---|---|---:|"
  echo "| \`UTCHMMA.2CTA\` | \`tcgen05.mma.cta_group::2\` (kind::tf32 and kind::f16/bf16; CTA pairs of the fused kernel) | $(grep -c 'UTCHMMA\.2CTA' /tmp/tfgnn_sass.txt) |"
  echo "| \`UTCHMMA\` (all) | \`tcgen05.mma\` | $(grep -c 'UTCHMMA' /tmp/tfgnn_sass.txt) |"
  echo "| \`UTMALDG\` | TMA tile loads (\`cp.async.bulk.tensor\`) | $(grep -c 'UTMALDG' /tmp/tfgnn_sass.txt) |"
  echo "| \`LDTM\` | \`tcgen05.ld\` (TMEM -> registers, epilogues) | $(grep -c 'LDTM' /tmp/tfgnn_sass.txt) |"
  echo "| \`UTCBAR\` | \`tcgen05.commit\` -> mbarrier | $(grep -c 'UTCBAR' /tmp/tfgnn_sass.txt) |"
  echo "| \`LDGSTS\` | \`cp.async\` (rolling gather ring of the fused kernel) | $(grep -c 'LDGSTS' /tmp/tfgnn_sass.txt) |"
  echo "| legacy \`HMMA\` / \`HGMMA\` (mma.sync / wgmma) | none expected | $(grep -cE '[^C]HMMA|HGMMA' /tmp/tfgnn_sass.txt) |"
  echo
  echo "tcgen05.mma instructions per kernel:"
  echo
  awk '/Function :/{fn=$3} /UTCHMMA/{c[fn]++} END{for(f in c) print c[f] " " f}' /tmp/tfgnn_sass.txt | sort -k2 | while read n f; do
    echo "- \`$(echo $f | c++filt | cut -c1-90)\`: $n"
  done
} > $OUT
cat $OUT | head -14
