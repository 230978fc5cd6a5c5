#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q -k "reference_newton_loop or reference_newmark_loop" 2>&1 | tail -25
