"""Test infrastructure: dumps the argparse surface of the reference's driver scripts (flags, defaults, types, actions)
into tests/golden/cli_flags.json.  Run in the build container, where /root/reference exists:

    python oracle/make_cli_golden.py
"""
import argparse
import json
import os
import re
import sys

REF = os.environ.get('VD_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'cli_flags.json')


def parser_of(script):
    src = open(os.path.join(REF, script)).read()
    block = src[src.index("parser = argparse.ArgumentParser"):src.index("args = parser.parse_args()")]
    block = '\n'.join(line[4:] if line.startswith('    ') else line for line in block.splitlines())
    ns = {'argparse': argparse}
    exec(block, ns)
    return ns['parser']


def dump(parser):
    out = {}
    for a in parser._actions:
        if not a.option_strings or a.dest == 'help':
            continue
        out[a.option_strings[0]] = {'default': a.default, 'type': getattr(a.type, '__name__', None),
                                    'action': type(a).__name__, 'choices': list(a.choices) if a.choices else None}
    return out


if __name__ == '__main__':
    data = {s: dump(parser_of(s)) for s in ('distill_s2d_ms.py', 'distill_baseline.py', 'buffer.py', 'distill_coreset.py')}
    json.dump(data, open(OUT, 'w'), indent=1, sort_keys=True)
    print('wrote', OUT, {k: len(v) for k, v in data.items()})
