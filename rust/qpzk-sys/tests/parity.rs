//! Cargo-side parity harness: identical seeded traces through the real
//! `PolynomialBatch::from_values` of qp-plonky2 1.1.1 and through `qpzk_batch_from_values`; the Merkle
//! caps, sampled LDE rows and Merkle paths must be equal. COMPILE-UNVERIFIED here (no Rust toolchain in
//! the build image); the same comparison runs against the C++ restatement in tests/test_gpu_parity.py.
use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::polynomial::PolynomialValues;
use plonky2::field::types::{Field, PrimeField64};
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::plonk::config::PoseidonGoldilocksConfig as C;
use plonky2::util::timing::TimingTree;
use qpzk_sys::*;

const P: u64 = 0xFFFF_FFFF_0000_0001;

/// SplitMix64 with rejection of values >= p (seed 0x5eed0001).
fn splitmix_trace(mut x: u64, n: usize) -> Vec<u64> {
    let mut out = Vec::with_capacity(n);
    while out.len() < n {
        x = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
        let mut z = x;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
        z ^= z >> 31;
        if z < P {
            out.push(z);
        }
    }
    out
}

#[test]
fn commit_matches_qp_plonky2() {
    let (k, ncols, rate_bits, cap_height) = (12usize, 135usize, 3usize, 4usize);
    let n = 1usize << k;
    let flat = splitmix_trace(0x5eed_0001, ncols * n);
    let values: Vec<PolynomialValues<F>> = flat
        .chunks(n)
        .map(|c| PolynomialValues::new(c.iter().map(|&v| F::from_canonical_u64(v)).collect()))
        .collect();
    let reference = PolynomialBatch::<F, C, 2>::from_values(
        values, rate_bits, false, cap_height, &mut TimingTree::default(), None);

    unsafe {
        let mut ctx = core::ptr::null_mut();
        assert_eq!(qpzk_ctx_create(0, 0, &mut ctx), QPZK_OK);
        let mut batch = core::ptr::null_mut();
        assert_eq!(
            qpzk_batch_from_values(ctx, flat.as_ptr(), ncols as u32, k as u32, rate_bits as u32,
                                   cap_height as u32, core::ptr::null(), 0, &mut batch),
            QPZK_OK
        );
        let mut cap = vec![0u64; 4 << cap_height];
        assert_eq!(qpzk_batch_cap(batch, cap.as_mut_ptr()), QPZK_OK);
        for (i, h) in reference.merkle_tree.cap.0.iter().enumerate() {
            for j in 0..4 {
                assert_eq!(h.elements[j].to_canonical_u64(), cap[4 * i + j]);
            }
        }
        let lde_len = n << rate_bits;
        let depth = k + rate_bits - cap_height;
        for s in 0..64u64 {
            let leaf = (s.wrapping_mul(0x9E37_79B9_7F4A_7C15) % lde_len as u64) as usize;
            let mut row = vec![0u64; ncols];
            let mut sib = vec![0u64; 4 * depth];
            assert_eq!(qpzk_batch_open(batch, leaf as u64, row.as_mut_ptr(), sib.as_mut_ptr()), QPZK_OK);
            let want_row = &reference.merkle_tree.leaves[leaf];
            assert!(want_row.iter().zip(&row).all(|(a, b)| a.to_canonical_u64() == *b));
            let proof = reference.merkle_tree.prove(leaf);
            for (l, h) in proof.siblings.iter().enumerate() {
                for j in 0..4 {
                    assert_eq!(h.elements[j].to_canonical_u64(), sib[4 * l + j]);
                }
            }
        }
        qpzk_batch_free(batch);
        qpzk_ctx_destroy(ctx);
    }
}

/// the byte layout of `write_polynomial_batch` is restated in libqpzk without a fixture (the
/// reference ships no serialized prover data). This is the pin: the bytes real qp-plonky2 writes for a batch must
/// equal `qpzk_batch_to_bytes` of the same commit, and `qpzk_batch_from_bytes` must accept them (with the
/// on-device verification) and serve the same cap.
#[test]
fn serialized_batch_matches_qp_plonky2() {
    use plonky2::util::serialization::{Buffer, Read, Write};
    let (k, ncols, rate_bits, cap_height) = (10usize, 84usize, 3usize, 4usize);
    let n = 1usize << k;
    let flat = splitmix_trace(0x5eed_0003, ncols * n);
    let values: Vec<PolynomialValues<F>> = flat
        .chunks(n)
        .map(|c| PolynomialValues::new(c.iter().map(|&v| F::from_canonical_u64(v)).collect()))
        .collect();
    let reference = PolynomialBatch::<F, C, 2>::from_values(
        values, rate_bits, false, cap_height, &mut TimingTree::default(), None);
    let mut want: Vec<u8> = Vec::new();
    want.write_polynomial_batch(&reference).unwrap();

    unsafe {
        let mut ctx = core::ptr::null_mut();
        assert_eq!(qpzk_ctx_create(0, 0, &mut ctx), QPZK_OK);
        let mut batch = core::ptr::null_mut();
        assert_eq!(
            qpzk_batch_from_values(ctx, flat.as_ptr(), ncols as u32, k as u32, rate_bits as u32,
                                   cap_height as u32, core::ptr::null(), 0, &mut batch),
            QPZK_OK
        );
        let mut len = 0u64;
        assert_eq!(qpzk_batch_serialized_size(batch, &mut len), QPZK_OK);
        assert_eq!(len as usize, want.len());
        let mut got = vec![0u8; len as usize];
        assert_eq!(qpzk_batch_to_bytes(batch, got.as_mut_ptr(), len), QPZK_OK);
        assert!(got == want, "serialized PolynomialBatch differs from qp-plonky2's");
        // and back: qp-plonky2 reads ours, we read qp-plonky2's
        let mut buf = Buffer::new(&got);
        let back: PolynomialBatch<F, C, 2> = buf.read_polynomial_batch().unwrap();
        assert_eq!(back.merkle_tree.cap, reference.merkle_tree.cap);
        let mut restored = core::ptr::null_mut();
        let mut used = 0u64;
        assert_eq!(
            qpzk_batch_from_bytes(ctx, want.as_ptr(), want.len() as u64, QPZK_IMPORT_VERIFY, &mut restored, &mut used),
            QPZK_OK
        );
        assert_eq!(used as usize, want.len());
        let mut cap = vec![0u64; 4 << cap_height];
        assert_eq!(qpzk_batch_cap(restored, cap.as_mut_ptr()), QPZK_OK);
        for (i, h) in reference.merkle_tree.cap.0.iter().enumerate() {
            for j in 0..4 {
                assert_eq!(h.elements[j].to_canonical_u64(), cap[4 * i + j]);
            }
        }
        qpzk_batch_free(restored);
        qpzk_batch_free(batch);
        qpzk_ctx_destroy(ctx);
    }
}
