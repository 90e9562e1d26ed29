'''Host-side mirror of the reference's plugin surface for the trace path (reference freecad_elements/).'''
