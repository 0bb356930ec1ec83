#!/bin/bash
# builds tuning variants of the library in parallel: name:flags[:ablations]   (ablations: nc nl nr nt)
build() { LG_LIB_SUFFIX=$1 LG_NVCC_EXTRA="$2" python -m loudgain_b200.build --force >/dev/null 2>&1 && echo "built $1"; }
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; flags=${rest%%:*}; abl=""
  [[ "$rest" == *:* ]] && abl=${rest#*:}
  build "$name" "$flags" &
  for a in $abl; do
    case $a in
      nc) build "${name}nc" "$flags -DLG_RUN_NOCOMP" & ;;
      nl) build "${name}nl" "$flags -DLG_RUN_NOLOAD=1" & ;;
      nr) build "${name}nr" "$flags -DLG_RUN_NOLOAD=2" & ;;
      nt) build "${name}nt" "$flags -DLG_RUN_NOTP" & ;;
    esac
  done
  wait
done
