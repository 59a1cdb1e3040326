//! uniffi export for the CUDA engine, the counterpart of `metal_msm_benchmark`
//! (example-app/src/lib.rs:15-26).  SOURCE ONLY: add next to the Metal export in example-app/src/lib.rs
//! (behind `#[cfg(feature = "cuda")]`, with `mopro-msm = { ..., features = ["cuda"] }` in example-app/Cargo.toml).
use mopro_msm::msm::cuda_msm::cuda_variable_base_msm;
use mopro_msm::msm::metal_msm::test_utils::generate_random_bases_and_scalars;

#[uniffi::export]
fn cuda_msm_benchmark(input_size: u32) -> () {
    let start = std::time::Instant::now();
    let (bases, scalars) = generate_random_bases_and_scalars(input_size as usize);
    println!("Generated bases and scalars in {:?}", start.elapsed());

    let start = std::time::Instant::now();
    let _result = cuda_variable_base_msm(&bases, &scalars).unwrap();
    println!("CUDA MSM took {:?}", start.elapsed());
}

#[cfg(test)]
mod tests {
    use super::*;
    use ark_bn254::G1Projective;
    use ark_ec::VariableBaseMSM;

    #[test]
    fn cuda_msm_matches_arkworks_2_16() {
        let (bases, scalars) = generate_random_bases_and_scalars(1 << 16);
        assert_eq!(cuda_variable_base_msm(&bases, &scalars).unwrap(), G1Projective::msm(&bases, &scalars).unwrap());
    }
}
