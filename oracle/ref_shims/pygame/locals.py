QUIT = 256
KEYDOWN = 768
KEYUP = 769
K_UP, K_DOWN, K_LEFT, K_RIGHT = 1073741906, 1073741905, 1073741904, 1073741903
K_w, K_s, K_a, K_d = 119, 115, 97, 100
__all__ = [n for n in dir() if n.isupper() or n.startswith("K_")]
