python -m pytest tests -q -m gpu 2>&1 | grep -E "^FAILED|passed|failed" | cut -c1-220 | head -40
