//! Drop-in for the gate-application hot path of `spinoza` (QuState/spinoza), backed by the B200 CUDA engine.
//!
//! Same names and argument order as the reference crate (`spinoza::core::State`, `spinoza::gates::{Gate, apply,
//! c_apply, cc_apply, mc_apply}`, `spinoza::circuit::QuantumCircuit`, `spinoza::measurement::measure_qubit`,
//! `spinoza::core::{iqft, qubit_expectation_value, xyz_expectation_value}`), so a user replaces
//! `use spinoza::...` with `use spinoza_b200::...`.
//!
//! Unavoidable deviation: the amplitudes live in GPU memory, so `state.reals` / `state.imags`
//! are methods that download (`state.reals()`), not `Vec` fields.  Where the reference panics
//! (`todo!()`, `unimplemented!()`, `assert!`) this crate panics too, with the engine's error text.
//!
//! This file has never been compiled in the repository's build environment (no Rust toolchain there); it is a
//! mechanical 1:1 binding of `include/spinoza_b200.h`.
#![allow(clippy::missing_safety_doc)]

use std::collections::HashSet;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

pub type Float = f64;
pub const PI: Float = std::f64::consts::PI;

pub mod ffi {
    use super::*;

    #[repr(C)]
    pub struct spz_state {
        _private: [u8; 0],
    }

    #[repr(C)]
    #[derive(Clone, Copy, Default)]
    pub struct spz_gate {
        pub kind: i32,
        pub t0: i32,
        pub t1: i32,
        pub reserved: i32,
        pub p: [f64; 3],
    }

    #[repr(C)]
    #[derive(Clone, Copy, Default)]
    pub struct spz_op {
        pub kind: i32,
        pub target: i32,
        pub t0: i32,
        pub t1: i32,
        pub p: [f64; 3],
        pub ctrl_kind: i32,
        pub reserved: i32,
        pub ctrl_mask: u64,
        pub zeros_mask: u64,
    }

    extern "C" {
        pub fn spz_last_error() -> *const c_char;
        pub fn spz_create(n_qubits: c_int, device: c_int, out: *mut *mut spz_state) -> c_int;
        pub fn spz_destroy(st: *mut spz_state) -> c_int;
        pub fn spz_clone(st: *const spz_state, out: *mut *mut spz_state) -> c_int;
        pub fn spz_num_qubits(st: *const spz_state) -> c_int;
        pub fn spz_len(st: *const spz_state) -> i64;
        pub fn spz_upload(st: *mut spz_state, re: *const f64, im: *const f64, offset: i64, count: i64) -> c_int;
        pub fn spz_download(st: *const spz_state, re: *mut f64, im: *mut f64, offset: i64, count: i64) -> c_int;
        pub fn spz_apply(st: *mut spz_state, gate: *const spz_gate, target: c_int) -> c_int;
        pub fn spz_c_apply(st: *mut spz_state, gate: *const spz_gate, control: c_int, target: c_int) -> c_int;
        pub fn spz_cc_apply(st: *mut spz_state, gate: *const spz_gate, c0: c_int, c1: c_int, target: c_int) -> c_int;
        pub fn spz_mc_apply_signed(st: *mut spz_state, gate: *const spz_gate, ones_mask: u64, zeros_mask: u64, target: c_int) -> c_int;
        pub fn spz_mc_apply(
            st: *mut spz_state,
            gate: *const spz_gate,
            controls: *const i32,
            n_controls: c_int,
            zeros: *const i32,
            n_zeros: c_int,
            target: c_int,
        ) -> c_int;
        pub fn spz_iqft(st: *mut spz_state, targets: *const i32, n_targets: c_int) -> c_int;
        pub fn spz_execute(
            st: *mut spz_state,
            ops: *const spz_op,
            n_ops: i64,
            flags: u32,
            measured_mask: *mut u64,
            measured_vals: *mut u64,
        ) -> c_int;
        pub fn spz_set_seed(st: *mut spz_state, seed: u64) -> c_int;
        pub fn spz_measure_qubit(st: *mut spz_state, target: c_int, reset: c_int, forced_v: c_int, out_bit: *mut c_int) -> c_int;
        pub fn spz_qubit_expectation_value(st: *mut spz_state, target: c_int, out: *mut f64) -> c_int;
        pub fn spz_xyz_expectation_value(
            st: *mut spz_state,
            observable: c_char,
            targets: *const i32,
            n_targets: c_int,
            out: *mut f64,
        ) -> c_int;
        pub fn spz_sample(st: *mut spz_state, u01: *const f64, shots: i64, out_index: *mut i64) -> c_int;
        pub fn spz_upload_async(st: *mut spz_state, re: *const f64, im: *const f64) -> c_int;
        pub fn spz_alloc_host(bytes: u64, out: *mut *mut std::ffi::c_void) -> c_int;
        pub fn spz_free_host(ptr: *mut std::ffi::c_void) -> c_int;
        pub fn spz_norm2(st: *mut spz_state, out: *mut f64) -> c_int;
        pub fn spz_sync(st: *mut spz_state) -> c_int;
        // sharded registers (include/spinoza_b200.h, "multi-GPU"): one process per GPU, or several shards in one process
        pub fn spz_dist_create(n_qubits: c_int, rank: c_int, world: c_int, device: c_int, out: *mut *mut spz_state) -> c_int;
        pub fn spz_dist_export(st: *mut spz_state, blob: *mut u8) -> c_int;
        pub fn spz_dist_connect(st: *mut spz_state, blobs: *const u8) -> c_int;
        pub fn spz_dist_connect_local(states: *mut *mut spz_state, world: c_int) -> c_int;
        pub fn spz_dist_connect_rdv(st: *mut spz_state, r: *mut spz_rdv) -> c_int;
        pub fn spz_dist_perm(st: *const spz_state, perm_out: *mut i32) -> c_int;
        pub fn spz_dist_local_qubits(st: *const spz_state) -> c_int;
        pub fn spz_dist_copy_from(dst: *mut spz_state, src: *const spz_state) -> c_int;
        pub fn spz_dist_stats(st: *const spz_state, out4: *mut f64) -> c_int;
        pub fn spz_rdv_open(dir: *const c_char, rank: c_int, world: c_int, out: *mut *mut spz_rdv) -> c_int;
        pub fn spz_rdv_allgather(r: *mut spz_rdv, mine: *const u8, bytes: i64, all: *mut u8) -> c_int;
        pub fn spz_rdv_barrier(r: *mut spz_rdv) -> c_int;
        pub fn spz_rdv_close(r: *mut spz_rdv) -> c_int;
    }

    #[repr(C)]
    pub struct spz_rdv {
        _private: [u8; 0],
    }
    pub const SPZ_IPC_BLOB_BYTES: usize = 256;
}

/// `spinoza::math` (math.rs): the amplitude type `Gate::to_matrix` returns.
pub mod math {
    use super::Float;

    #[derive(Clone, Copy, Debug, PartialEq)]
    pub struct Amplitude {
        pub re: Float,
        pub im: Float,
    }
}

fn check(status: c_int) {
    if status != 0 {
        let msg = unsafe { CStr::from_ptr(ffi::spz_last_error()) }.to_string_lossy().into_owned();
        panic!("spinoza-b200 status {status}: {msg}"); // the reference panics in the same places
    }
}

pub mod core {
    use super::*;

    /// `spinoza::core::State` (core.rs:18-51), device-resident.
    pub struct State {
        pub(crate) h: *mut ffi::spz_state,
        pub n: u8,
    }
    unsafe impl Send for State {}

    impl State {
        /// core.rs:32-42
        pub fn new(n: usize) -> Self {
            assert!(n > 0);
            let mut h = std::ptr::null_mut();
            check(unsafe { ffi::spz_create(n as c_int, 0, &mut h) });
            Self { h, n: n as u8 }
        }
        /// Build from host vectors (the reference lets tests construct `State { reals, imags, n }` directly).
        pub fn from_vecs(reals: &[Float], imags: &[Float]) -> Self {
            assert_eq!(reals.len(), imags.len());
            let n = reals.len().trailing_zeros() as usize;
            let s = Self::new(n);
            check(unsafe { ffi::spz_upload(s.h, reals.as_ptr(), imags.as_ptr(), 0, reals.len() as i64) });
            s
        }
        #[allow(clippy::len_without_is_empty)]
        pub fn len(&self) -> usize {
            unsafe { ffi::spz_len(self.h) as usize }
        }
        /// `state.reals` of the reference (download).
        pub fn reals(&self) -> Vec<Float> {
            let mut v = vec![0.0; self.len()];
            check(unsafe { ffi::spz_download(self.h, v.as_mut_ptr(), std::ptr::null_mut(), 0, v.len() as i64) });
            v
        }
        /// `state.imags` of the reference (download).
        pub fn imags(&self) -> Vec<Float> {
            let mut v = vec![0.0; self.len()];
            check(unsafe { ffi::spz_download(self.h, std::ptr::null_mut(), v.as_mut_ptr(), 0, v.len() as i64) });
            v
        }
        pub fn amp(&self, i: usize) -> (Float, Float) {
            let (mut re, mut im) = (0.0, 0.0);
            check(unsafe { ffi::spz_download(self.h, &mut re, &mut im, i as i64, 1) });
            (re, im)
        }
        pub fn set_seed(&mut self, seed: u64) {
            check(unsafe { ffi::spz_set_seed(self.h, seed) });
        }
        /// Whole-state upload that returns at once; the gates issued next follow the state piece by piece as it arrives.
        /// # Safety
        /// `reals` / `imags` must be page-locked (`spz_alloc_host`), hold `len()` values each and stay alive and untouched until
        /// the next `sync` / `reals()` / `imags()`.
        pub unsafe fn upload_async(&mut self, reals: *const Float, imags: *const Float) {
            check(ffi::spz_upload_async(self.h, reals, imags));
        }
        pub fn sync(&mut self) {
            check(unsafe { ffi::spz_sync(self.h) });
        }
    }
    impl Clone for State {
        fn clone(&self) -> Self {
            let mut h = std::ptr::null_mut();
            check(unsafe { ffi::spz_clone(self.h, &mut h) });
            Self { h, n: self.n }
        }
    }
    impl Drop for State {
        fn drop(&mut self) {
            unsafe { ffi::spz_destroy(self.h) };
        }
    }

    /// core.rs:184-191
    pub fn iqft(state: &mut State, targets: &[usize]) {
        let t: Vec<i32> = targets.iter().map(|&x| x as i32).collect();
        check(unsafe { ffi::spz_iqft(state.h, t.as_ptr(), t.len() as c_int) });
    }
    /// core.rs:198-219
    pub fn qubit_expectation_value(state: &State, target: usize) -> Float {
        let mut out = 0.0;
        check(unsafe { ffi::spz_qubit_expectation_value(state.h, target as c_int, &mut out) });
        out
    }
    /// core.rs:222-264
    pub fn xyz_expectation_value(observable: char, state: &State, targets: &[usize]) -> Vec<Float> {
        let t: Vec<i32> = targets.iter().map(|&x| x as i32).collect();
        let mut out = vec![0.0; t.len()];
        check(unsafe {
            ffi::spz_xyz_expectation_value(state.h, observable as c_char, t.as_ptr(), t.len() as c_int, out.as_mut_ptr())
        });
        out
    }
    /// `spinoza::core::Reservoir` (core.rs:65-121).  Same construction and read-out; the filling is the engine's exact
    /// inverse-CDF sampler (one read pass over the device state, `spz_sample`) instead of `num_tests` host-side weighted
    /// replacement rounds, so every entry is an exact draw from |amplitude|^2 whatever `num_tests` is.
    pub struct Reservoir {
        entries: Vec<usize>,
        seed: u64,
    }

    impl Reservoir {
        /// core.rs:73-78
        pub fn new(k: usize) -> Self {
            let t = std::time::SystemTime::now().duration_since(std::time::UNIX_EPOCH).map(|d| d.as_nanos() as u64).unwrap_or(0);
            Self { entries: vec![0; k], seed: t ^ 0x9E37_79B9_7F4A_7C15 } // the reference draws from thread_rng: unseeded
        }
        /// Deterministic draws (tests).
        pub fn with_seed(k: usize, seed: u64) -> Self {
            Self { entries: vec![0; k], seed }
        }
        fn uniforms(&mut self) -> Vec<Float> {
            // splitmix64, 53 mantissa bits per draw: the generator the engine and its oracle use everywhere
            let mut x = self.seed;
            let u: Vec<Float> = (0..self.entries.len())
                .map(|_| {
                    x = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
                    let mut z = x;
                    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
                    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
                    ((z ^ (z >> 31)) >> 11) as Float * (1.0 / 9007199254740992.0)
                })
                .collect();
            self.seed = x;
            u
        }
        /// core.rs:101-112 takes the host slices of the state; here the state lives on the device, so it takes the `State`.
        /// `num_tests` is accepted for source compatibility (see the type's comment).
        pub fn sampling(&mut self, state: &State, _num_tests: usize) {
            let u = self.uniforms();
            self.entries = sample(state, &u).into_iter().map(|i| i as usize).collect();
        }
        /// core.rs:115-121
        pub fn get_outcome_count(&self) -> std::collections::HashMap<usize, usize> {
            let mut samples = std::collections::HashMap::new();
            for e in &self.entries {
                *samples.entry(*e).or_insert(0) += 1;
            }
            samples
        }
    }

    /// core.rs:125-129
    pub fn reservoir_sampling(state: &State, reservoir_size: usize, num_tests: usize) -> Reservoir {
        let mut reservoir = Reservoir::new(reservoir_size);
        reservoir.sampling(state, num_tests);
        reservoir
    }

    /// Exact inverse-CDF sampling (what `Reservoir::sampling` runs on, core.rs:125): one basis-state index per uniform.
    pub fn sample(state: &State, u01: &[Float]) -> Vec<i64> {
        let mut out = vec![0i64; u01.len()];
        check(unsafe { ffi::spz_sample(state.h, u01.as_ptr(), u01.len() as i64, out.as_mut_ptr()) });
        out
    }
}

pub mod gates {
    use super::core::State;
    use super::*;

    /// gates.rs:44-74 (`Unitary` is out of scope for the B200 engine)
    #[derive(Clone, Debug, PartialEq)]
    pub enum Gate {
        H,
        M,
        X,
        Y,
        Z,
        P(Float),
        RX(Float),
        RY(Float),
        RZ(Float),
        SWAP(usize, usize),
        U(Float, Float, Float),
        BitFlipNoise(Float),
    }

    impl Gate {
        /// gates.rs:78-92
        pub fn inverse(self) -> Self {
            match self {
                Self::H | Self::X | Self::Y | Self::Z | Self::SWAP(_, _) => self,
                Self::P(t) => Self::P(-t),
                Self::RX(t) => Self::RX(-t),
                Self::RY(t) => Self::RY(-t),
                Self::RZ(t) => Self::RZ(-t),
                Self::U(theta, phi, lambda) => Self::U(-theta, -lambda, -phi),
                Self::M | Self::BitFlipNoise(_) => unimplemented!(),
            }
        }
        /// gates.rs:95-190: the 2 x 2 matrix, row-major.  (M, SWAP, BitFlipNoise: `unimplemented!()` like the reference.)
        pub fn to_matrix(&self) -> [math::Amplitude; 4] {
            let a = |re: Float, im: Float| math::Amplitude { re, im };
            let r = std::f64::consts::FRAC_1_SQRT_2;
            match *self {
                Self::H => [a(r, 0.0), a(r, 0.0), a(r, 0.0), a(-r, 0.0)],
                Self::X => [a(0.0, 0.0), a(1.0, 0.0), a(1.0, 0.0), a(0.0, 0.0)],
                Self::Y => [a(0.0, 0.0), a(0.0, -1.0), a(0.0, 1.0), a(0.0, 0.0)],
                Self::Z => [a(1.0, 0.0), a(0.0, 0.0), a(0.0, 0.0), a(-1.0, 0.0)],
                Self::P(t) => [a(1.0, 0.0), a(0.0, 0.0), a(0.0, 0.0), a(t.cos(), t.sin())],
                Self::RX(t) => {
                    let (s, c) = (t / 2.0).sin_cos();
                    [a(c, 0.0), a(0.0, -s), a(0.0, -s), a(c, 0.0)]
                }
                Self::RY(t) => {
                    let (s, c) = (t / 2.0).sin_cos();
                    [a(c, 0.0), a(-s, 0.0), a(s, 0.0), a(c, 0.0)]
                }
                Self::RZ(t) => {
                    let (s, c) = (t / 2.0).sin_cos();
                    [a(c, -s), a(0.0, 0.0), a(0.0, 0.0), a(c, s)]
                }
                Self::U(theta, phi, lambda) => {
                    let (s, c) = (theta / 2.0).sin_cos();
                    [
                        a(c, 0.0),
                        a(-lambda.cos() * s, -lambda.sin() * s),
                        a(phi.cos() * s, phi.sin() * s),
                        a((phi + lambda).cos() * c, (phi + lambda).sin() * c),
                    ]
                }
                _ => unimplemented!(),
            }
        }
        pub(crate) fn to_ffi(&self) -> ffi::spz_gate {
            let mut g = ffi::spz_gate::default();
            match *self {
                Self::H => g.kind = 0,
                Self::M => g.kind = 1,
                Self::X => g.kind = 2,
                Self::Y => g.kind = 3,
                Self::Z => g.kind = 4,
                Self::P(t) => { g.kind = 5; g.p[0] = t }
                Self::RX(t) => { g.kind = 6; g.p[0] = t }
                Self::RY(t) => { g.kind = 7; g.p[0] = t }
                Self::RZ(t) => { g.kind = 8; g.p[0] = t }
                Self::SWAP(a, b) => { g.kind = 9; g.t0 = a as i32; g.t1 = b as i32 }
                Self::U(a, b, c) => { g.kind = 10; g.p = [a, b, c] }
                Self::BitFlipNoise(p) => { g.kind = 12; g.p[0] = p }
            }
            g
        }
    }

    /// gates.rs:215
    pub fn apply(gate: Gate, state: &mut State, target: usize) {
        check(unsafe { ffi::spz_apply(state.h, &gate.to_ffi(), target as c_int) });
    }
    /// gates.rs:257
    pub fn c_apply(gate: Gate, state: &mut State, control: usize, target: usize) {
        check(unsafe { ffi::spz_c_apply(state.h, &gate.to_ffi(), control as c_int, target as c_int) });
    }
    /// gates.rs:272
    pub fn cc_apply(gate: Gate, state: &mut State, control0: usize, control1: usize, target: usize) {
        check(unsafe { ffi::spz_cc_apply(state.h, &gate.to_ffi(), control0 as c_int, control1 as c_int, target as c_int) });
    }
    /// Extension (not in the reference): every qubit of `ones` must be 1 and every qubit of `zeros` must be 0 -- the negative
    /// controls `Controls::Mixed { zeros }` describes and `mc_apply` drops (gates.rs:298-311).
    pub fn mc_apply_signed(gate: Gate, state: &mut State, ones: &[usize], zeros: &[usize], target: usize) {
        let om = ones.iter().fold(0u64, |m, &q| m | (1u64 << q));
        let zm = zeros.iter().fold(0u64, |m, &q| m | (1u64 << q));
        check(unsafe { ffi::spz_mc_apply_signed(state.h, &gate.to_ffi(), om, zm, target as c_int) });
    }
    /// gates.rs:290
    pub fn mc_apply(gate: Gate, state: &mut State, controls: &[usize], zeros: Option<HashSet<usize>>, target: usize) {
        let c: Vec<i32> = controls.iter().map(|&x| x as i32).collect();
        let z: Vec<i32> = zeros.as_ref().map(|s| s.iter().map(|&x| x as i32).collect()).unwrap_or_default();
        check(unsafe {
            ffi::spz_mc_apply(
                state.h,
                &gate.to_ffi(),
                c.as_ptr(),
                c.len() as c_int,
                if zeros.is_some() { z.as_ptr() } else { std::ptr::null() },
                z.len() as c_int,
                target as c_int,
            )
        });
    }
}

pub mod measurement {
    use super::core::State;
    use super::*;

    /// measurement.rs:12
    pub fn measure_qubit(state: &mut State, target: usize, reset: bool, v: Option<u8>) -> u8 {
        let mut bit: c_int = 0;
        let forced = v.map(|x| x as c_int).unwrap_or(-1);
        check(unsafe { ffi::spz_measure_qubit(state.h, target as c_int, reset as c_int, forced, &mut bit) });
        bit as u8
    }
}

pub mod circuit {
    use super::core::State;
    use super::gates::Gate;
    use super::*;

    /// circuit.rs:13-51
    #[derive(Clone)]
    pub struct QuantumRegister(pub Vec<usize>);
    impl std::ops::Index<usize> for QuantumRegister {
        type Output = usize;
        fn index(&self, i: usize) -> &usize {
            &self.0[i]
        }
    }
    impl QuantumRegister {
        pub fn new(size: usize) -> Self {
            assert!(size > 0);
            QuantumRegister((0..size).collect())
        }
        #[allow(clippy::len_without_is_empty)]
        pub fn len(&self) -> usize {
            self.0.len()
        }
        pub fn update_shift(&mut self, shift: usize) {
            self.0.iter_mut().for_each(|x| *x += shift);
        }
        pub fn get_shift(&self) -> usize {
            self[0]
        }
    }

    /// circuit.rs:55-109
    #[derive(Clone)]
    pub enum Controls {
        None,
        Single(usize),
        Ones(Vec<usize>),
        Mixed { controls: Vec<usize>, zeros: HashSet<usize> },
    }
    impl Controls {
        fn from(controls: &[usize], zeros: Option<HashSet<usize>>) -> Self {
            if let Some(zs) = zeros {
                Self::Mixed { controls: controls.to_vec(), zeros: zs }
            } else if controls.is_empty() {
                Self::None
            } else if controls.len() == 1 {
                Self::Single(controls[0])
            } else {
                Self::Ones(controls.to_vec())
            }
        }
        fn unpack(&self) -> (Vec<usize>, HashSet<usize>) {
            match self {
                Self::None => (vec![], HashSet::new()),
                Self::Single(c) => (vec![*c], HashSet::new()),
                Self::Ones(cs) => (cs.clone(), HashSet::new()),
                Self::Mixed { controls, zeros } => (controls.clone(), zeros.clone()),
            }
        }
        fn new_with_control(&self, control: usize, shift: usize) -> Self {
            let (mut controls, zeros) = self.unpack();
            controls.iter_mut().for_each(|c| *c += shift);
            controls.push(control);
            if zeros.is_empty() {
                Self::from(&controls, None)
            } else {
                Self::from(&controls, Some(zeros.iter().map(|z| z + shift).collect()))
            }
        }
    }

    /// circuit.rs:113-120
    #[derive(Clone)]
    pub struct QuantumTransformation {
        pub gate: Gate,
        pub target: usize,
        pub controls: Controls,
    }

    /// circuit.rs:168-601.  `execute` hands the whole list to the engine's fusing scheduler.
    pub struct QuantumCircuit {
        pub transformations: Vec<QuantumTransformation>,
        pub state: State,
        measured_qubits: u64,      // QubitTracker circuit.rs:122-164
        measured_qubits_vals: u64,
        pub quantum_registers_info: Vec<usize>,
        /// SPZ_EXEC_* flags; default = fused
        pub exec_flags: u32,
    }

    impl QuantumCircuit {
        pub fn new(registers: &mut [&mut QuantumRegister]) -> Self {
            let mut bits = 0;
            let mut sizes = Vec::with_capacity(registers.len());
            for r in registers.iter_mut() {
                r.update_shift(bits);
                sizes.push(r.len());
                bits += r.len();
            }
            Self::from_state(State::new(bits), sizes)
        }
        pub fn from_state(state: State, quantum_registers_info: Vec<usize>) -> Self {
            Self { transformations: Vec::new(), state, measured_qubits: 0, measured_qubits_vals: 0, quantum_registers_info, exec_flags: 1 }
        }
        pub fn get_statevector(&self) -> &State {
            &self.state
        }
        pub fn inverse(&mut self) {
            self.transformations.reverse();
            self.transformations.iter_mut().for_each(|qt| qt.gate = qt.gate.clone().inverse());
        }
        #[inline]
        pub fn add(&mut self, t: QuantumTransformation) {
            self.transformations.push(t);
        }
        fn g(&mut self, gate: Gate, target: usize, controls: Controls) {
            self.add(QuantumTransformation { gate, target, controls });
        }
        pub fn measure(&mut self, t: usize) { self.g(Gate::M, t, Controls::None) }
        pub fn swap(&mut self, t0: usize, t1: usize) { self.g(Gate::SWAP(t0, t1), 0, Controls::None) }
        pub fn x(&mut self, t: usize) { self.g(Gate::X, t, Controls::None) }
        pub fn y(&mut self, t: usize) { self.g(Gate::Y, t, Controls::None) }
        pub fn z(&mut self, t: usize) { self.g(Gate::Z, t, Controls::None) }
        pub fn h(&mut self, t: usize) { self.g(Gate::H, t, Controls::None) }
        pub fn p(&mut self, a: Float, t: usize) { self.g(Gate::P(a), t, Controls::None) }
        pub fn rx(&mut self, a: Float, t: usize) { self.g(Gate::RX(a), t, Controls::None) }
        pub fn ry(&mut self, a: Float, t: usize) { self.g(Gate::RY(a), t, Controls::None) }
        pub fn rz(&mut self, a: Float, t: usize) { self.g(Gate::RZ(a), t, Controls::None) }
        pub fn u(&mut self, th: Float, ph: Float, la: Float, t: usize) { self.g(Gate::U(th, ph, la), t, Controls::None) }
        pub fn cx(&mut self, c: usize, t: usize) { self.g(Gate::X, t, Controls::Single(c)) }
        pub fn ccx(&mut self, c1: usize, c2: usize, t: usize) { self.g(Gate::X, t, Controls::Ones(vec![c1, c2])) }
        pub fn ch(&mut self, c: usize, t: usize) { self.g(Gate::H, t, Controls::Single(c)) }
        pub fn cy(&mut self, c: usize, t: usize) { self.g(Gate::Y, t, Controls::Single(c)) }
        pub fn cp(&mut self, a: Float, c: usize, t: usize) { self.g(Gate::P(a), t, Controls::Single(c)) }
        pub fn crx(&mut self, a: Float, c: usize, t: usize) { self.g(Gate::RX(a), t, Controls::Single(c)) }
        pub fn cry(&mut self, a: Float, c: usize, t: usize) { self.g(Gate::RY(a), t, Controls::Single(c)) }
        pub fn crz(&mut self, a: Float, c: usize, t: usize) { self.g(Gate::RZ(a), t, Controls::Single(c)) }
        pub fn cu(&mut self, th: Float, ph: Float, la: Float, c: usize, t: usize) { self.g(Gate::U(th, ph, la), t, Controls::Single(c)) }
        pub fn bit_flip_noise(&mut self, prob: Float, t: usize) { self.g(Gate::BitFlipNoise(prob), t, Controls::None) }
        /// circuit.rs:438-445
        pub fn iqft(&mut self, targets: &[usize]) {
            for j in (0..targets.len()).rev() {
                self.h(targets[j]);
                for k in (0..j).rev() {
                    self.cp(-PI / (2.0 as Float).powi((j - k) as i32), targets[j], targets[k]);
                }
            }
        }
        pub fn append(&mut self, circuit: &QuantumCircuit, reg: &QuantumRegister) {
            assert_eq!(reg.len(), circuit.quantum_registers_info.iter().sum::<usize>());
            for tr in circuit.transformations.iter() {
                self.g(tr.gate.clone(), reg.get_shift() + tr.target, tr.controls.clone());
            }
        }
        pub fn c_append(&mut self, circuit: &QuantumCircuit, c: usize, reg: &QuantumRegister) {
            assert!(!(reg.get_shift()..reg.get_shift() + reg.len()).contains(&c));
            for tr in circuit.transformations.iter() {
                self.g(tr.gate.clone(), reg.get_shift() + tr.target, tr.controls.new_with_control(c, reg.get_shift()));
            }
        }
        pub fn mc_append(&mut self, circuit: &QuantumCircuit, controls: &[usize], reg: &QuantumRegister) {
            let range = reg.get_shift()..reg.get_shift() + reg.len();
            for c in controls {
                assert!(!range.contains(c), "control {c} should not be in: Range(start: {} end: {})", range.start, range.end);
            }
            for control in controls {
                for tr in circuit.transformations.iter() {
                    self.g(tr.gate.clone(), reg.get_shift() + tr.target, tr.controls.new_with_control(*control, reg.get_shift()));
                }
            }
        }
        pub fn is_qubit_measured(&self, q: usize) -> bool {
            ((self.measured_qubits >> q) & 1) == 1
        }
        pub fn get_qubit_measured_val(&self, q: usize) -> Option<u8> {
            self.is_qubit_measured(q).then(|| ((self.measured_qubits_vals >> q) & 1) as u8)
        }
        /// circuit.rs:552-600
        pub fn execute(&mut self) {
            let ops: Vec<ffi::spz_op> = self
                .transformations
                .drain(..)
                .map(|tr| {
                    let g = tr.gate.to_ffi();
                    let mut op = ffi::spz_op { kind: g.kind, target: tr.target as i32, t0: g.t0, t1: g.t1, p: g.p, ..Default::default() };
                    match &tr.controls {
                        Controls::None => op.ctrl_kind = 0,
                        Controls::Single(c) => { op.ctrl_kind = 1; op.ctrl_mask = 1u64 << c }
                        Controls::Ones(cs) => { op.ctrl_kind = 2; cs.iter().for_each(|c| op.ctrl_mask |= 1u64 << c) }
                        Controls::Mixed { controls, zeros } => {
                            op.ctrl_kind = 3;
                            controls.iter().for_each(|c| op.ctrl_mask |= 1u64 << c);
                            zeros.iter().for_each(|z| op.zeros_mask |= 1u64 << z);
                        }
                    }
                    op
                })
                .collect();
            check(unsafe {
                ffi::spz_execute(self.state.h, ops.as_ptr(), ops.len() as i64, self.exec_flags, &mut self.measured_qubits, &mut self.measured_qubits_vals)
            });
        }
    }
}

/// Sharded registers: the 2^n amplitudes spread over the GPUs of one node by the top log2(world) index bits
/// (include/spinoza_b200.h, "multi-GPU"; no counterpart in the reference).  A shard IS a `core::State`: `gates::apply`,
/// `circuit::QuantumCircuit::execute`, `measurement::measure_qubit` and the reductions are the same calls, made by every rank
/// in the same order; gates on a global qubit exchange half a shard with the partner GPU over NVLink inside the call.
pub mod dist {
    use super::core::State;
    use super::*;

    /// The host-side control plane between the processes of a node (csrc/rendezvous.cu): files under /dev/shm, no framework.
    pub struct Rendezvous {
        h: *mut ffi::spz_rdv,
        pub rank: usize,
        pub world: usize,
    }
    impl Rendezvous {
        /// `dir = None`: one directory per launch, derived from MASTER_PORT and the launcher's pid.
        pub fn open(dir: Option<&str>, rank: usize, world: usize) -> Self {
            let c = dir.map(|d| std::ffi::CString::new(d).expect("rendezvous directory"));
            let mut h = std::ptr::null_mut();
            check(unsafe { ffi::spz_rdv_open(c.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()), rank as c_int, world as c_int, &mut h) });
            Self { h, rank, world }
        }
        /// RANK / WORLD_SIZE as torchrun- or mpirun-style launchers export them.
        pub fn from_env() -> Self {
            let get = |k: &str, d: usize| std::env::var(k).ok().and_then(|v| v.parse().ok()).unwrap_or(d);
            Self::open(std::env::var("SPZ_RDV_DIR").ok().as_deref(), get("RANK", 0), get("WORLD_SIZE", 1))
        }
        pub fn barrier(&mut self) {
            check(unsafe { ffi::spz_rdv_barrier(self.h) });
        }
        /// Equal-length blobs, one per rank, in rank order.
        pub fn all_gather(&mut self, mine: &[u8]) -> Vec<u8> {
            let mut all = vec![0u8; mine.len() * self.world];
            check(unsafe { ffi::spz_rdv_allgather(self.h, mine.as_ptr(), mine.len() as i64, all.as_mut_ptr()) });
            all
        }
        pub fn max_float(&mut self, x: Float) -> Float {
            self.all_gather(&x.to_le_bytes()).chunks(8).map(|c| Float::from_le_bytes(c.try_into().unwrap())).fold(Float::MIN, Float::max)
        }
    }
    impl Drop for Rendezvous {
        fn drop(&mut self) {
            unsafe { ffi::spz_rdv_close(self.h) };
        }
    }

    /// This rank's shard of an n-qubit register, |0..0>, connected to the other ranks' shards (CUDA IPC over NVLink).
    pub fn sharded_state(n: usize, rdv: &mut Rendezvous, device: i32) -> State {
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::spz_dist_create(n as c_int, rdv.rank as c_int, rdv.world as c_int, device as c_int, &mut h) });
        check(unsafe { ffi::spz_dist_connect_rdv(h, rdv.h) });
        State { h, n: n as u8 }
    }

    /// All `world` shards in ONE process (tests, or one process driving several GPUs): plain device pointers, same kernels.
    pub fn local_group(n: usize, world: usize, devices: &[i32]) -> Vec<State> {
        let mut hs: Vec<*mut ffi::spz_state> = (0..world)
            .map(|r| {
                let mut h = std::ptr::null_mut();
                check(unsafe { ffi::spz_dist_create(n as c_int, r as c_int, world as c_int, devices[r % devices.len()] as c_int, &mut h) });
                h
            })
            .collect();
        check(unsafe { ffi::spz_dist_connect_local(hs.as_mut_ptr(), world as c_int) });
        hs.into_iter().map(|h| State { h, n: n as u8 }).collect()
    }

    /// Where each logical qubit lives now (physical index bit; bits >= `local_qubits` are rank bits).
    pub fn perm(state: &State) -> Vec<usize> {
        let mut p = vec![0i32; state.n as usize];
        check(unsafe { ffi::spz_dist_perm(state.h, p.as_mut_ptr()) });
        p.into_iter().map(|x| x as usize).collect()
    }
    pub fn local_qubits(state: &State) -> usize {
        unsafe { ffi::spz_dist_local_qubits(state.h) as usize }
    }
    /// Collective copy of a sharded register (`Clone` of the reference's `State`, rank by rank).
    pub fn copy_from(dst: &mut State, src: &State) {
        check(unsafe { ffi::spz_dist_copy_from(dst.h, src.h) });
    }
    pub struct Stats {
        pub exchanges: f64,
        pub bytes_sent: f64,
        pub exchange_ms: f64,
        pub overlapped: f64,
    }
    pub fn stats(state: &State) -> Stats {
        let mut o = [0.0f64; 4];
        check(unsafe { ffi::spz_dist_stats(state.h, o.as_mut_ptr()) });
        Stats { exchanges: o[0], bytes_sent: o[1], exchange_ms: o[2], overlapped: o[3] }
    }
    /// sum |amplitude|^2 over all ranks (a collective).
    pub fn norm2(state: &State) -> Float {
        let mut out = 0.0;
        check(unsafe { ffi::spz_norm2(state.h, &mut out) });
        out
    }
}
