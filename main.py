"""Command-line entry with the reference's flag surface (main.py:14-41,50 of lanl/pyDNMFk) for the update-loop path.

    python main.py --p_r=1 --p_c=1 --k=4 --fpath=data/ --fname=swim --ftype=mat --init=rand --itr=1000 --norm=fro --method=mu
    torchrun --nproc-per-node 4 main.py --p_r=4 --p_c=1 ...         (one process per GPU instead of mpirun -n 4)

`--process=pyDNMFk` runs NMFk (perturbation ensemble, clustering, silhouettes, regression fit, rank selection) for
k = start_k .. end_k and prints the estimated rank."""
import argparse
import sys

from pydnmfk_b200 import config
from pydnmfk_b200.dist_comm import MPI, MPI_comm
from pydnmfk_b200.data_io import data_read
from pydnmfk_b200.pyDNMF import PyNMF
from pydnmfk_b200.pyDNMFk import PyNMFk
from pydnmfk_b200.utils import str2bool


# Flag surface of the reference CLI (names, types and defaults must match for scripts to keep working).
_NMF_FLAGS = [
    ('p_r', int, None, 'processor-grid rows'),
    ('p_c', int, None, 'processor-grid columns'),
    ('k', int, 4, 'factorization rank'),
    ('fpath', str, 'data/', 'directory of the input'),
    ('ftype', str, 'mat', 'input format: mat / npy / csv / folder'),
    ('fname', str, 'A_', 'input file stem'),
    ('init', str, 'rand', 'factor initialisation: rand / nnsvd (nnsvd: 1-D grids only, like the reference)'),
    ('itr', int, 5000, 'update iterations'),
    ('norm', str, 'kl', 'objective: kl or fro'),
    ('method', str, 'mu', 'update rule: mu, hals or bcd'),
    ('verbose', str2bool, False, 'print the relative error'),
    ('results_path', str, 'results/', 'output directory'),
    ('checkpoint', str2bool, False, 'NMFk checkpointing'),
    ('timing_stats', str2bool, False, 'collect per-function timings'),
    ('prune', str2bool, False, 'drop all-zero rows / columns before factorizing'),
    ('precision', str, 'float32', 'float32 or float64'),
]
_NMFK_FLAGS = [
    ('perturbations', int, 20, 'ensemble size per k'),
    ('noise_var', float, 0.015, 'perturbation amplitude'),
    ('start_k', int, 1, 'first rank of the sweep'),
    ('end_k', int, 10, 'last rank of the sweep'),
    ('step_k', int, 1, 'rank increment'),
    ('sill_thr', float, 0.6, 'silhouette threshold of the rank selection'),
    ('sampling', str, 'uniform', 'perturbation law: uniform or poisson'),
]


def _add(parser, table):
    for name, typ, default, text in table:
        if default is None:
            parser.add_argument('--' + name, type=typ, required=True, help=text)
        else:
            parser.add_argument('--' + name, type=typ, default=default, help=text)
    return parser


def parser_pyNMF(parser):
    return _add(parser, _NMF_FLAGS)


def parser_pyNMFk(parser):
    return _add(parser, _NMFK_FLAGS)


def main(argv=None):
    parser = argparse.ArgumentParser(description='Arguments for pyDNMF/pyDNMFk')
    parser.add_argument('--process', type=str, default='pyDNMF', help='pyDNMF/pyDNMFk')
    parser = parser_pyNMF(parser)
    parser = parser_pyNMFk(parser)
    try:
        args = parser.parse_args(argv)
    except SystemExit:
        parser.print_help()
        sys.exit(0)
    config.flag = args.timing_stats
    main_comm = MPI.COMM_WORLD
    rank = main_comm.rank
    comm = MPI_comm(main_comm, args.p_r, args.p_c)
    args.rank = rank
    args.comm1 = comm.comm
    args.comm = comm
    args.col_comm = comm.cart_1d_column()
    args.row_comm = comm.cart_1d_row()
    if rank == 0:
        print('Reading data now')
    A_ij = data_read(args).read()
    if rank == 0:
        print('Reading data complete')
    if args.process == 'pyDNMFk':
        if rank == 0:
            print('Starting PyDNMFk...')
        out = PyNMFk(A_ij, factors=None, params=args).fit()
        if rank == 0:
            print('PyDNMFk done.')
    else:
        if rank == 0:
            print('Starting PyDNMF...')
        out = PyNMF(A_ij, factors=None, params=args).fit()
        if rank == 0:
            print('PyDNMF done. relative error = %s' % out[2])
    if rank == 0 and args.timing_stats:
        # reference main.py:83-88: print the per-function timings and write them as a one-row table
        # (pandas.DataFrame([config.time]).to_csv layout); the plot of plot_results.py is out of scope
        print(config.time)
        keys = list(config.time.keys())
        with open(args.results_path + 'Timing_stats.csv', 'w') as f:
            f.write(',' + ','.join(keys) + '\n0,' + ','.join(str(config.time[key]) for key in keys) + '\n')
    return out


if __name__ == '__main__':
    main()
