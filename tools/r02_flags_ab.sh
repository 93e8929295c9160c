#!/bin/bash
# In-call A/B of megakernel flag words (CRABML_MEGA_FLAGS): phase profile + device-resident tok/s for each.
#   usage (GPU box): tools/r02_flags_ab.sh "0x1 0x1805 ..." [Q8_0|Q4_0]
FL=${1:-"0x1 0x5 0x9 0x1005 0x1805 0x2005 0x180d"}; WL=${2:-Q8_0}
for f in $FL; do
  echo "== flags $f ($WL)"
  CRABML_MEGA_FLAGS=$f timeout 150 python tools/mega_profile.py $WL 2>&1 | grep -E "tokens back|token total|n= |rror" | head -12
done
