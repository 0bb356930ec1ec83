# Whole-wave tuning of the packed sweep: resident CTAs per SM (32-thread CTAs, variant t32) x chunks per slot
run() { echo "variant=[$1] ctas=$2 k=$3"; LG_LIB_SUFFIX=$1 LOUDGAIN_B200_PAIR_CTAS=$2 LOUDGAIN_B200_CHUNKS_PER_SLOT=$3 python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-300; }
run "" 0 5
run t32 0 5
run t32 16 5
run t32 15 5
run t32 14 5
run t32 14 7
run t32 12 6
run t32 10 5
run t32 15 10
run t32 13 5
