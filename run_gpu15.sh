#!/bin/bash
timeout 120 python scripts/debug_wgrad.py 2>&1 | tail -20
