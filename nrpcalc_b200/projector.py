"""Scratch project directory of one Finder job (mirrors reference nrpcalc/base/projector.py:1-19).

finder() keeps two side files in ./<uuid>/: seq_list.txt (written by the graph builder, re-read for the
output dict) and repeat_graph.txt (adjacency-list dump used to restore the graph after each vertex-cover
round).  The directory is removed at the end of the job and, as a safety net, at interpreter exit.
"""
import atexit
import os
import shutil

current_uuid = None


def setup_proj_dir(uuid):
    global current_uuid
    current_uuid = uuid
    directory = './{}/'.format(current_uuid)
    if not os.path.exists(directory):
        os.makedirs(directory)


@atexit.register
def remove_proj_dir():
    global current_uuid
    if current_uuid is None:
        return
    directory = './{}/'.format(current_uuid)
    if os.path.isdir(directory):
        shutil.rmtree(directory)
