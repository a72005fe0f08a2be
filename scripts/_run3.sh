cd /root/repo
for v in 0 1 2 3 4 5; do echo "variant $v"; HNR_CHAIN_VARIANT=$v python scripts/bench_chain.py 2>&1 | grep "chain"; done
