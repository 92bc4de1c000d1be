echo "== base =="; python tools/_variants/d64_perf.py
for v in emu1 emu2 emu3 emu4; do echo "== $v =="; FA_SM100_LIB=tools/_variants/lib_$v.so python tools/_variants/d64_perf.py; done
echo "== base again =="; python tools/_variants/d64_perf.py
