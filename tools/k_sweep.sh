for k in 1 2 3 5; do echo "k=$k"; LOUDGAIN_B200_CHUNKS_PER_SLOT=$k python bench.py --quick --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-250; done
