set -x
timeout 300 python scripts/debug_gram.py 2>&1 | head -300
