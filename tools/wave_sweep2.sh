run() { echo "fmt=$1 k=$2"; LOUDGAIN_B200_CHUNKS_PER_SLOT=$2 python bench.py --quick --format $1 --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-300; }
run s16 0
run s16 6
run s16 7
run s16 9
run f32 0
run f32 7
