for d in 0 1 2 4 6 7; do EEGAN_TS_DBG=$d timeout 100 python scratch/mainloop_probe2.py; done
