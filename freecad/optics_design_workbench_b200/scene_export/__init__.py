'''Scene export: FreeCAD document / FCStd archive -> flat, immutable device scene description.'''
