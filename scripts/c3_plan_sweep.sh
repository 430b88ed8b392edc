# c3 kernel time against the band width (CVGS_TMA_NPB, 32 columns each) and the ring depth (CVGS_TMA_SLOTS)
for npb in ${NPBS:-4 3 2 1}; do for sl in ${SLOTS:-2 3}; do
  echo -n "NPB=$npb SLOTS=$sl  "; CVGS_TMA_NPB=$npb CVGS_TMA_SLOTS=$sl python scripts/run_case.py c3 30 0 2>&1 | grep "us/launch"
done; done
