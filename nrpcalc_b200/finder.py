"""Finder-mode orchestrator -- host wrapper with the behaviour of reference nrpcalc/base/finder.py:15-196:
argument checks, scratch project directory, graph build (the GPU hot path), independent-set recovery, output
dictionary in seq_list.txt order and the optional FASTA file."""
import textwrap
import uuid

from . import hgraph
from . import kmerSetDB
from . import projector
from . import recovery


def _check_finder_constraints(seq_list, homology, allow_internal_repeat, vercov_func):
    for seq in seq_list:
        if not isinstance(seq, str):
            print(' [ERROR]    Parts in Sequence List must be string, not {}'.format(type(seq)))
            print(' [SOLUTION] Try correcting Lmax\n')
            return False
    if not isinstance(homology, int):
        print('\n [ERROR]    Lmax must be an integer, not {}'.format(type(homology)))
        print(' [SOLUTION] Try correcting Lmax\n')
        return False
    if homology - 1 < 5:
        print('\n [ERROR]    Lmax must be greater than 4, not {}'.format(homology - 1))
        print(' [SOLUTION] Try correcting Lmax\n')
        return False
    if allow_internal_repeat not in [True, False]:
        print('\n [ERROR]    Internal Repeat must be boolean, not {}'.format(type(allow_internal_repeat)))
        print(' [SOLUTION] Try correcting Internal Repeat\n')
        return False
    return True


def _check_finder_arguments(vercov_func, output_file):
    if not isinstance(vercov_func, str):
        print(' [ERROR]    Vertex Cover Routine Name must be string, not {}'.format(type(vercov_func)))
        print(' [SOLUTION] Try correcting Vertex Cover Routine Name\n')
        return False
    if vercov_func not in ['2apx', 'nrpG', 'nrp2']:
        print('\n [ERROR]    Vertex Cover Routine Name must be \'2apx\', \'nrpG\', or \'nrp2\', not \'{}\''.format(vercov_func))
        print(' [SOLUTION] Try correcting Vertex Cover Routine Name\n')
        return False
    if output_file is not None and not isinstance(output_file, str):
        print('\n [ERROR]    Output File must be a string or None, not {}'.format(type(output_file)))
        print(' [SOLUTION] Try correcting Output File\n')
        return False
    return True


def nrp_finder(seq_list, homology, allow_internal_repeat=False, background=None, vercov_func=None, output_file=None,
               verbose=True, check_constraints=True):
    if verbose:
        print('\n[Non-Repetitive Parts Calculator - Finder Mode]')
    seq_list = list(seq_list)          # the caller's list is never consumed (finder.py:82)
    ok = True

    if check_constraints:
        if verbose:
            print('\n[Checking Constraints]')
            print(' Sequence List   : {} parts'.format(len(seq_list)))
            print('          Lmax   : {} bp'.format(homology - 1))
            print(' Internal Repeats: {}'.format(allow_internal_repeat))
        ok = _check_finder_constraints(seq_list, homology, allow_internal_repeat, vercov_func)
        if verbose:
            print(' Check Status: FAIL\n' if not ok else '\n Check Status: PASS')

        if ok and background is not None:
            if verbose:
                print('\n[Checking Background]\n Background: {}'.format(background))
            if isinstance(background, kmerSetDB.kmerSetDB):
                if not background.ALIVE:
                    # the reference reports this but does not abort here (finder.py:111-116 sets an unused name);
                    # the RuntimeError surfaces later from the membership test
                    print('\n [ERROR]    Background is closed or dropped')
                    print(' [SOLUTION] Try using an open Background\n')
                    if verbose:
                        print(' Check Status: FAIL\n')
                elif verbose:
                    print('\n Check Status: PASS')
            else:
                ok = False
                print('\n [ERROR]    Background Object is INVALID')
                print(' [SOLUTION] Try instantiating background via nrpcalc.background(...)\n')
                if verbose:
                    print(' Check Status : FAIL\n')

        if ok:
            if verbose:
                print('\n[Checking Arguments]')
                print('   Vertex Cover: {}'.format(vercov_func))
                print('   Output  File: {}'.format(output_file))
            ok = _check_finder_arguments(vercov_func, output_file)
            if verbose:
                print(' Check Status: FAIL\n' if not ok else '\n Check Status: PASS')

        if not ok:
            raise RuntimeError('Invalid Constraints, or Background')

    if verbose:
        print()

    current_uuid = str(uuid.uuid4())
    projector.setup_proj_dir(current_uuid)
    seq_file = './{}/seq_list.txt'.format(current_uuid)
    graph_file = './{}/repeat_graph.txt'.format(current_uuid)

    try:
        homology_graph = hgraph.get_homology_graph(seq_list, background, homology, allow_internal_repeat, seq_file, verbose)
        indi_seqs_indx = recovery.get_recovered_non_homologs(homology_graph, graph_file, vercov_func, verbose)
    except BaseException:
        projector.remove_proj_dir()        # the reference leaves ./<uuid>/ behind when the build raises; the error itself is unchanged
        raise

    parts_dict = {}
    with open(seq_file, 'r') as infile:
        for line in infile:
            index, seq = line.strip().split(',')
            index = int(index)
            if index in indi_seqs_indx:
                parts_dict[index] = seq

    if output_file:
        with open(output_file, 'w') as outfile:
            for index, seq in parts_dict.items():
                outfile.write('>index {}\n'.format(index))
                outfile.write('{}\n'.format('\n'.join(textwrap.wrap(seq, 80))))

    projector.remove_proj_dir()
    if verbose:
        print('\nNon-Repetitive Toolbox Size: {}'.format(len(parts_dict)))
    return parts_dict
