"""In-tree locations of the native artefacts."""
import os

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libmerging_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG), "include")
