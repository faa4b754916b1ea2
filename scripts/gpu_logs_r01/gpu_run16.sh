timeout 600 python -m pytest tests -x -q -m gpu -k "device_cut" 2>&1 | grep -v "^cut nfree" | tail -3
python scripts/cut_bench.py 2>&1 | grep -E "^cut nfree|device batch" | awk '/^cut nfree/{c++; if(c%12==1) print; next} {print}' | head -30
