"""Opt-in layout options of the union-tile kernels through the C ABI (layout only: the kernels follow the step words):
MFT_OPT_TILE bit 4 (tuned second record copy) and MFT_OPT_REFINE_ORDER (rows of a tile ordered by D' row length) give the
same results as the default layout.  (Written without GPU access: first hardware run is the round-end pass.)"""
import pytest

from test_zz_ee_edge_sizes_gpu import _case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("source", ["upwind", "residual"])
def test_tuned_second_copy_layout(source):
    """MFT_OPT_TILE = 31: the second record copy's bank groups come from the builder's local search (layout only; the kernels
    follow the step words) -- same results as every other layout"""
    import mft_b200 as m

    _case(m, 40, 36, 20, 31, source)


@pytest.mark.parametrize("tile", [15, 31, 0])
def test_rows_of_a_tile_ordered_by_transposed_row_length(tile):
    """MFT_OPT_REFINE_ORDER: a permutation inside every tile (256-row blocks for the sliced-ELL kernels) -- same results"""
    import mft_b200 as m

    _case(m, 40, 36, 20, tile, "residual", refine_order=True)


@pytest.mark.parametrize("tile,tile_rows", [(15, 21), (31, 21), (31, 22), (31, 42)])
def test_rows_per_thread_with_the_tuned_layout(tile, tile_rows):
    """two / four rows per thread over the union of their stencils (pass B digit first), default and tuned record copies"""
    import mft_b200 as m

    _case(m, 40, 36, 20, tile, "residual", tile_rows=tile_rows)
