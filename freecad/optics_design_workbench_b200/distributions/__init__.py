'''Tabulated source sampler + fan grid (host side); see sampler_tables.py.'''
from .sampler_tables import *
