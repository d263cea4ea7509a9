# what the driver runs at round end: the GPU test suite, smoke(), a short default bench
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
