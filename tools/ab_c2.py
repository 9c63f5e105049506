"""Same-box A/B of library builds on the C2 sweep: python tools/ab_c2.py libA.so libB.so ...  (each library in
its own process, alternating, 2 rounds).  DEVELOPMENT TOOL."""
import os
import subprocess
import sys
libs = sys.argv[1:]
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ, LINS='5e-2')
        if lib != 'default':
            env['STB_LIB'] = os.path.abspath(lib)
        out = subprocess.run([sys.executable, 'tools/prof_lin.py'], env=env, capture_output=True, text=True).stdout
        line = [l for l in out.splitlines() if l.startswith('lin_rms')]
        print('%-28s %s' % (lib, line[-1][:100] if line else out[-300:]), flush=True)
