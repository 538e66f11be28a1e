"""Identification strings reported over GTP and in tree dumps (program.py:3-37 of the reference)."""
PROGRAM_NAME = "TamaGo"
PROTOCOL_VERSION = "2"
VERSION = "0.10.0"
