debug = False
