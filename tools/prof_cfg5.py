import cProfile, pstats, os, sys, io
sys.argv = ['bench_cfg5.py', '--quick1000']
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import runpy
rank = int(os.environ.get('RANK', '0'))
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'bench_cfg5.py'), run_name='__main__')
except SystemExit:
    pass
finally:
    pr.disable()
    if rank == 0:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
        print(s.getvalue()[:9000])
