'''Simulation-side host code: scene/source setup, hit writer in the reference's on-disk format.'''
