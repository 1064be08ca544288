#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-160
