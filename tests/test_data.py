"""CPU tests of the data ingest (pgmvae/data.py; reference run.py:52-56): whole-file parse, chunked streaming parse."""
import os

import numpy as np
import pytest

from pgmvae import data


@pytest.mark.parametrize("trailing_newline", [True, False])
@pytest.mark.parametrize("rows_per_chunk", [1, 7, 300, 5000])
def test_chunked_csv_reader_reproduces_the_file(tmp_path, trailing_newline, rows_per_chunk):
    y = data.synthetic_binary(1000, 9, seed=1)
    p = tmp_path / "toy.data"
    p.write_text("\n".join(",".join(str(int(t)) for t in r) for r in y) + ("\n" if trailing_newline else ""))
    chunks = list(data.iter_binary_csv(str(p), rows_per_chunk, nvar=9))
    assert all(c.dtype == np.uint8 and c.shape[1] == 9 and 0 < len(c) <= rows_per_chunk for c in chunks)
    np.testing.assert_array_equal(np.concatenate(chunks), y)
    np.testing.assert_array_equal(data.parse_binary_csv(str(p), 9), y)


def test_chunked_reader_on_the_shipped_dataset_layout(tmp_path):
    """nltcs (cfg1) as run.py reads it: the packed copy shipped with the package, written back as the reference's CSV."""
    y = data.load_split("nltcs", "valid", 16)
    p = tmp_path / "nltcs.valid.data"
    p.write_text("".join(",".join(str(int(t)) for t in r) + "\n" for r in y))
    np.testing.assert_array_equal(np.concatenate(list(data.iter_binary_csv(str(p), 512, 16))), y)
    with pytest.raises(ValueError):
        list(data.iter_binary_csv(str(p), 512, 17))


def test_iter_array_covers_everything():
    y = data.synthetic_binary(103, 5)
    np.testing.assert_array_equal(np.concatenate(list(data.iter_array(y, 10))), y)
