for q in 32768 16384 8192 4096; do echo "blocks_per_cta=$q"; LOUDGAIN_B200_QUERY_BLOCKS_PER_CTA=$q python bench.py --quick --steps 30 --warmup 3 2>&1 | tail -1 | cut -c1-200; done
