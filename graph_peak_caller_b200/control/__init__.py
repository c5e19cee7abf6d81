"""Background (lambda) track: the functions ``CallPeaks`` calls
(reference graph_peak_caller/control/__init__.py:7-42), same names and
arguments.  The window sizes are the only thing the two front ends differ in;
the work is SparseControl.create (kernels gpc_control_linear_track and
gpc_control_backproject)."""
import logging

from .controlgenerator import SparseControl
from .linearmap import LinearMap  # noqa: F401  (re-exported as in the reference)

# linear windows the read starts are spread over; None stands for the fragment length
WINDOWS_WITH_CONTROL = (None, 1000, 10000)
WINDOWS_FROM_INPUT = (10000,)


def get_background_track(graph, intervals, config, extensions, touched_nodes=None):
    control = SparseControl(config.linear_map_name, graph, extensions, config.fragment_length,
                            touched_nodes)
    floor = config.global_min
    if floor is not None:
        control.set_min_value(floor)
    return control.create(intervals)


def _windows(config, spec):
    return [config.fragment_length if w is None else w for w in spec]


def get_background_track_from_control(graph, intervals, config, touched_nodes=None):
    logging.info("Background from control reads")
    return get_background_track(graph, intervals, config,
                                _windows(config, WINDOWS_WITH_CONTROL), touched_nodes)


def get_background_track_from_input(graph, intervals, config, touched_nodes=None):
    logging.info("Background from the input reads themselves")
    return get_background_track(graph, intervals, config,
                                _windows(config, WINDOWS_FROM_INPUT), touched_nodes)


def scale_tracks(fragment_pileup, background_track, ratio):
    """Brings the track of the larger read set down to the smaller one, in place
    (lazily for device tracks): ratio = sample reads / control reads."""
    if ratio == 1:
        return
    if ratio < 1:
        logging.info("Control scaled by %.3f" % ratio)
        background_track *= ratio
        return
    logging.warning("More reads in sample than in control")
    fragment_pileup *= 1 / ratio
