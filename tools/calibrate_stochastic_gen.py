import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import ref_pool
if __name__ == "__main__":
    t=time.time()
    S = ref_pool.run(ref_pool.fmr_spectra, 12288, 8192, 128)
    print("gen", time.time()-t, S.shape)
    np.save('/tmp/fmr_pool.npy', S)
    ref_pool.close()
