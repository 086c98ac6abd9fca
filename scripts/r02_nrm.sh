set -x
python scripts/variants.py run cfg5
python scripts/variants.py run cfg2
