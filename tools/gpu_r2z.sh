#!/bin/bash
python -c "
import json, bench
print(json.dumps(bench.table_section(0), indent=1))
" 2>&1 | tail -40
