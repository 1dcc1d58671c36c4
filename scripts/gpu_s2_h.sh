set -x
timeout 300 python scripts/debug_gram.py 2>&1 | grep -v "^--" | head -30
timeout 200 python -m pytest tests/test_gpu_units.py -m gpu -q --timeout=120 -p no:cacheprovider -k "gram" 2>&1 | grep -v "^$" | tail -30
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --maxfail=15 -p no:cacheprovider -k "not gram" 2>&1 | tail -15
