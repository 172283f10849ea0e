for d in 0 1 2 4 8 32 7 15 47; do echo "debug=$d"; OFQ_ATTN_DEBUG=$d python tools/attn_bench.py 2>&1 | grep "codes + out only\|+ qp16$" ; done
