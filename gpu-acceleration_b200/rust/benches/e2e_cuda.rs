//! Criterion arm for the CUDA engine, beside the reference's `metal_msm` / `arkworks_msm` arms
//! (mopro-msm/benches/e2e.rs:22-64: same group settings, same sizes, datasets generated outside the timed closure,
//! `Throughput::Elements(n)`).  SOURCE ONLY: this image has no cargo/rustc.
//!
//! Add to mopro-msm/Cargo.toml:
//!     [[bench]]
//!     name = "e2e_cuda"
//!     harness = false
//!     required-features = ["cuda"]
//! Run:  cargo bench --features cuda --bench e2e_cuda
use ark_bn254::G1Projective as G;
use ark_ec::VariableBaseMSM;
use criterion::{criterion_group, criterion_main, BenchmarkId, Criterion, Throughput};
use mopro_msm::msm::cuda_msm::{cuda_variable_base_msm, RegisteredBases};
use mopro_msm::msm::metal_msm::test_utils::generate_random_bases_and_scalars;
use std::time::Duration;

// the reference's sizes (benches/e2e.rs: 2^10, 2^12, 2^16) plus BASELINE.json's 2^20 / 2^24
const LOG_SIZES: &[usize] = &[10, 12, 16, 20, 24];

fn bench_e2e_cuda(c: &mut Criterion) {
    let mut group = c.benchmark_group("msm_e2e");
    group.measurement_time(Duration::from_secs(2));
    group.sample_size(10);
    for &log_n in LOG_SIZES {
        let n = 1usize << log_n;
        let (bases, scalars) = generate_random_bases_and_scalars(n);
        group.throughput(Throughput::Elements(n as u64));

        // the drop-in call: host slices in, one point out (uploads inside the timed closure)
        group.bench_with_input(BenchmarkId::new("cuda_msm", n), &n, |b, &_n| {
            b.iter(|| {
                let _res = cuda_variable_base_msm(&bases, &scalars).unwrap();
            });
        });

        // the proving-key pattern: bases registered once, only the scalars move
        let key = RegisteredBases::new(&bases, true).unwrap();
        group.bench_with_input(BenchmarkId::new("cuda_msm_registered", n), &n, |b, &_n| {
            b.iter(|| {
                let _res = key.msm(&scalars).unwrap();
            });
        });

        // the CPU baseline, exactly as in the reference's bench
        if log_n <= 20 {
            group.bench_with_input(BenchmarkId::new("arkworks_msm", n), &n, |b, &_n| {
                b.iter(|| {
                    let _res = G::msm(&bases, &scalars).unwrap();
                });
            });
        }
    }
    group.finish();
}

criterion_group!(benches, bench_e2e_cuda);
criterion_main!(benches);
