timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== fixed"; timeout 300 python scripts/kb_mel.py 2>&1 | tail -1
echo "== generic"; IRIS_NO_FIXED_EPI=1 timeout 300 python scripts/kb_mel.py 2>&1 | tail -1
