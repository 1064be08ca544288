#!/bin/bash
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -1; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
