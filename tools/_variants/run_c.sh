echo "== persistent =="; python tools/quick_perf.py --no-parity --short
echo "== one item per CTA =="; FA_SM100_PERSISTENT=0 python tools/quick_perf.py --no-parity --short
echo "== margin 8 =="; FA_SM100_SM_MARGIN=8 python tools/quick_perf.py --no-parity --short
