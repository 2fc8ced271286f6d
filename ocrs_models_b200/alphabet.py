"""The recognition alphabet the reference trains with (ocrs_models/datasets/hiertext.py:133-137): 96 distinct
characters -> 97 classes with the CTC blank at index 0 (labels are 1-based, datasets/util.py:113-129)."""

DEFAULT_ALPHABET = (
    " 0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~" + chr(8364) + "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
)
assert len(DEFAULT_ALPHABET) == 96 and len(set(DEFAULT_ALPHABET)) == 96
