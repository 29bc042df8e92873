python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
python bench.py 2>&1 | tail -1 | cut -c1-600
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
