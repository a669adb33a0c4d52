// Replacement for the body of the render thread in render() (src/render/mod.rs:984-1026).  Signature, the cancel-watcher
// thread (:947-958), the PPM writer (:1031-1088) and RenderDone stay exactly as they are; the progress thread (:965-982) is
// replaced by the preview callback below, which sends the same RenderUpdate{progress, image} the GUI draws
// (src/main.rs:340-397, src/views/render_tab.rs:278-297) -- with a real partial image instead of the half-filled pixel list.
// UNTESTED IN THIS REPO (no cargo here); the C side of every call below is exercised by tests/ through ctypes.
//
//     let render_thread_handle = s.spawn(move || {
//         let scene = &render_config.scene;
//         println!("Rendering scene {} ({} objects), {} samples per pixel, {}x{} resolution (B200 backend)", ...);
//
            let (objs, tris, cam) = ffi::flatten(scene);
            let desc = ffi::ptb_scene_desc { objects: objs.as_ptr(), n_objects: objs.len() as u64,
                                             triangles: tris.as_ptr(), n_triangles: tris.len() as u64, camera: cam };
            // every GPU of the box in ONE context: samples per pixel are split across them inside ptb_render*
            let n_gpus = unsafe { ffi::ptb_device_count() };
            assert!(n_gpus > 0, "no B200: the backend has no CPU fallback");
            let ids: Vec<i32> = (0..n_gpus).collect();
            let mut ctx: *mut ffi::ptb_ctx = std::ptr::null_mut();
            unsafe {
                assert_eq!(ffi::ptb_create_multi(ids.as_ptr(), n_gpus, &mut ctx), ffi::PTB_OK);
                assert_eq!(ffi::ptb_upload_scene(ctx, &desc), ffi::PTB_OK);
            }
            // stop_render is polled through `cancel` (a watcher copies the AtomicBool into it every 100 ms, omitted);
            // processed_pixel_count's role is taken by `samples_done`
            let cancel = std::sync::atomic::AtomicI32::new(0);
            let samples_done = std::sync::atomic::AtomicU64::new(0);
            let mut out = vec![0f32; grid_size * 3];
            let spp = render_config.samples_per_pixel as u64;
            let seed: u64 = rand::random();                       // the reference is OS-seeded too (mod.rs:53)

            // RenderUpdate every 500 ms: the library resolves the partial sum and calls back on this thread
            struct PreviewCtx<'a> { sink: &'a mut Sink<RenderUpdate>, res: Resolution }
            extern "C" fn on_preview(user: *mut std::ffi::c_void, mean_rgb: *const f32, w: i32, h: i32, spp_done: u64, spp_total: u64) {
                let p = unsafe { &mut *(user as *mut PreviewCtx) };
                let px = unsafe { std::slice::from_raw_parts(mean_rgb, (w * h * 3) as usize) };
                let pixels: Vec<Vec3> = px.chunks_exact(3).map(|c| Vec3::new(c[0], c[1], c[2])).collect();
                let _ = futures::executor::block_on(p.sink.send(RenderUpdate {
                    progress: spp_done as f32 / spp_total as f32,
                    image: Image::new(pixels, p.res),             // same buffer order as the rayon path (mod.rs:805-806)
                }));
            }
            let mut pctx = PreviewCtx { sink: update_sink, res };
            let rc = unsafe {
                ffi::ptb_render_progressive(ctx, res.width as i32, res.height as i32, 0, spp, seed, ffi::PTB_OUT_MEAN, out.as_mut_ptr(),
                                            cancel.as_ptr() as *const i32, samples_done.as_ptr(), 500.0, Some(on_preview),
                                            &mut pctx as *mut _ as *mut std::ffi::c_void)
            };
            assert!(rc >= 0, "ptb_render_progressive failed");
            {   // final image: index i <-> x = i % W, y = H-1 - i / W (mod.rs:805-806)
                let mut px = pixels.lock().unwrap();
                for (i, p) in px.iter_mut().enumerate() { *p = Vec3::new(out[3 * i], out[3 * i + 1], out[3 * i + 2]); }
            }
            unsafe { ffi::ptb_destroy(ctx) };
//         println!("Rendering complete");
//         stop_render.store(true, atomic::Ordering::Relaxed);
//         ... PPM writer and Image::new unchanged ...
